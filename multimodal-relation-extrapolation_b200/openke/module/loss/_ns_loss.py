"""The three negative-sampling losses of OpenKE (OpenKE/openke/module/loss/{MarginLoss,SigmoidLoss,SoftplusLoss}.py) over
ONE library kernel: mre_ns_loss computes the loss value and dLoss/dscore of the strategy's [B, 1] / [B, neg] score blocks in a
single launch (self-adversarial weights included), so a loss module here is a thin autograd node, not a chain of eager
elementwise ops.  There is no CPU path: the score blocks must live on the GPU.
"""
import torch
import torch.nn as nn

from .... import _lib as L
from .... import engine


def _flat_scores(p_score, n_score):
    """[B, 1] positives and [B, neg] negatives -> the strategy's flat layout [B | neg blocks of B] (contiguous float32)"""
    B = p_score.shape[0]
    if p_score.numel() != B or n_score.dim() != 2 or n_score.shape[0] != B:
        raise ValueError(f"expected p_score [B, 1] and n_score [B, neg], got {tuple(p_score.shape)} and {tuple(n_score.shape)}")
    return torch.cat([p_score.reshape(-1), n_score.t().reshape(-1)]).to(torch.float32).contiguous(), B, n_score.shape[1]


class _NSLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p_score, n_score, kind, margin, adv, temperature, lib_ctx):
        if not p_score.is_cuda:
            raise L.MreError("mre_b200 losses compute on a B200 only (mre_ns_loss); the scores are on " + str(p_score.device))
        flat, B, neg = _flat_scores(p_score.detach(), n_score.detach())
        loss = torch.empty(1, dtype=torch.float32, device=flat.device)
        dscore = torch.empty_like(flat)
        L.check(L.lib().mre_ns_loss(lib_ctx._h, kind, flat.data_ptr(), B, neg, float(margin), int(adv), float(temperature),
                                    loss.data_ptr(), dscore.data_ptr(), torch.cuda.current_stream().cuda_stream))
        ctx.save_for_backward(dscore)
        ctx.shape = (B, neg)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dscore,) = ctx.saved_tensors
        B, neg = ctx.shape
        d = dscore * g
        return d[:B].view(B, 1), d[B:].view(neg, B).t(), None, None, None, None, None


class NegativeSamplingLoss(nn.Module):
    """Common part: the (frozen) hyper-parameters under the reference's state_dict names, the library context of the
    scores' device, and forward() -> loss tensor of shape [1] as the reference's modules return."""

    kind = None

    def __init__(self, adv_temperature=None, margin=None):
        super().__init__()
        frozen = lambda v: nn.Parameter(torch.tensor([float(v)]), requires_grad=False)
        self.zero_const, self.pi_const = frozen(0.0), frozen(3.14159265358979323846)          # BaseModule.py:7-12
        if margin is not None:
            self.margin = frozen(margin)
        self.adv_flag = adv_temperature is not None
        if self.adv_flag:
            self.adv_temperature = frozen(adv_temperature)
        self._ctx = {}

    def lib_ctx(self, device):
        idx = device.index or 0
        if idx not in self._ctx:
            self._ctx[idx] = engine.Context(idx)
        return self._ctx[idx]

    def hyper(self):
        """(margin, adv flag, temperature) as plain numbers for the C ABI"""
        return (float(self.margin.item()) if hasattr(self, "margin") else 0.0, int(self.adv_flag),
                float(self.adv_temperature.item()) if self.adv_flag else 0.0)

    def adv_sign(self):
        return 1.0

    def get_weights(self, n_score):
        """the detached self-adversarial weights of a row's negatives (softmax over the last axis)"""
        return torch.softmax(self.adv_sign() * n_score * self.adv_temperature, dim=-1).detach()

    def forward(self, p_score, n_score):
        margin, adv, temp = self.hyper()
        return _NSLossFn.apply(p_score, n_score, self.kind, margin, adv, temp, self.lib_ctx(p_score.device))

    def predict(self, p_score, n_score):
        return self.forward(p_score, n_score).cpu().data.numpy()
