"""SoftplusLoss (OpenKE/openke/module/loss/SoftplusLoss.py:7-33): (mean softplus(-p) + mean_b sum_k w_bk softplus(n_bk)) / 2,
w = 1/neg or softmax_k(T n_bk) -- kind MRE_LOSS_SOFTPLUS of mre_ns_loss."""
from .... import _lib as L
from ._ns_loss import NegativeSamplingLoss


class SoftplusLoss(NegativeSamplingLoss):
    kind = L.LOSS_SOFTPLUS

    def __init__(self, adv_temperature=None):
        super().__init__(adv_temperature=adv_temperature)
