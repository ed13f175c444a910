"""SigmoidLoss (OpenKE/openke/module/loss/SigmoidLoss.py:7-32): -(mean logsigmoid(p) + mean_b sum_k w_bk logsigmoid(-n_bk)) / 2,
w = 1/neg or softmax_k(T n_bk) -- kind MRE_LOSS_SIGMOID of mre_ns_loss."""
from .... import _lib as L
from ._ns_loss import NegativeSamplingLoss


class SigmoidLoss(NegativeSamplingLoss):
    kind = L.LOSS_SIGMOID

    def __init__(self, adv_temperature=None):
        super().__init__(adv_temperature=adv_temperature)
