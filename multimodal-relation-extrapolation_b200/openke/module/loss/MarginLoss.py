"""MarginLoss (OpenKE/openke/module/loss/MarginLoss.py:10-32; the paper's copy module/loss.py:5-28):
mean_b sum_k w_bk max(p_b - n_bk, -margin) + margin, w = 1/neg or softmax_k(-T n_bk) -- kind MRE_LOSS_MARGIN of mre_ns_loss."""
from .... import _lib as L
from ._ns_loss import NegativeSamplingLoss


class MarginLoss(NegativeSamplingLoss):
    kind = L.LOSS_MARGIN

    def __init__(self, adv_temperature=None, margin=6.0):
        super().__init__(adv_temperature=adv_temperature, margin=margin)

    def adv_sign(self):
        return -1.0           # lower distance = harder negative (MarginLoss.py:21-22)
