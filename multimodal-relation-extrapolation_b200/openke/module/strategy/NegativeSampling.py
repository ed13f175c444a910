"""NegativeSampling strategy (OpenKE/openke/module/strategy/NegativeSampling.py:5-32): splits the flat score vector
into p_score [B, 1] and n_score [B, neg] and applies the loss.  When nothing but the loss feeds the gradient (no regulariser:
the configuration of OpenKE/examples/train_transe_FB15K237.py:23-39, BASELINE configs[3]) `fused_step` runs the whole
forward + loss + backward in mre_ns_train_step and leaves the gradients in .grad, so Trainer skips autograd; otherwise
forward() is differentiable end to end (library score and loss kernels behind autograd nodes)."""
import torch
import torch.nn as nn

from .... import engine
from ..model.Model import expand_batch


class NegativeSampling(nn.Module):
    def __init__(self, model=None, loss=None, batch_size=256, regul_rate=0.0, l3_regul_rate=0.0):
        super().__init__()
        self.model = model
        self.loss = loss
        self.batch_size = batch_size
        self.regul_rate = regul_rate
        self.l3_regul_rate = l3_regul_rate

    def _get_positive_score(self, score):
        positive_score = score[:self.batch_size]
        return positive_score.view(-1, self.batch_size).permute(1, 0)

    def _get_negative_score(self, score):
        negative_score = score[self.batch_size:]
        return negative_score.view(-1, self.batch_size).permute(1, 0)

    def forward(self, data):
        score = self.model(data)
        p_score = self._get_positive_score(score)
        n_score = self._get_negative_score(score)
        loss_res = self.loss(p_score, n_score)
        if self.regul_rate != 0:
            loss_res += self.regul_rate * self.model.regularization(data)
        if self.l3_regul_rate != 0:
            loss_res += self.l3_regul_rate * self.model.l3_regularization()
        return loss_res

    # ---- fused path
    def can_fuse(self):
        """one mre_ns_train_step (score -> loss -> backward, three launches) replaces forward + autograd when the loss is one of
        the library's and nothing else feeds the gradient (no regulariser, no margin-shifted score, single-pass scorer)"""
        from ..loss._ns_loss import NegativeSamplingLoss
        from ..model.SimplE import SimplE
        return (isinstance(self.loss, NegativeSamplingLoss) and getattr(self.model, "scorer", None) in engine.SCORERS
                and not isinstance(self.model, SimplE) and getattr(self.model, "fusable", True)
                and not getattr(self.model, "margin_flag", False)
                and self.regul_rate == 0 and self.l3_regul_rate == 0)

    def fused_step(self, data):
        """loss [1] of one batch with d loss / d tables ACCUMULATED into the embeddings' .grad (as backward() would)"""
        m = self.model
        tabs = m.tables()
        h, t, r = expand_batch(data, tabs[0].device)
        n = h.numel()
        B = self.batch_size
        assert n % B == 0 and n > B
        for w in tabs:
            if w.grad is None:
                w.grad = torch.zeros_like(w)
        margin, adv, temp = self.loss.hyper()
        kw = m.rank_kwargs()
        return engine.ns_train_step(m.ctx(), m.scorer, [w.data for w in tabs], [w.grad for w in tabs], h, t, r, B, n // B - 1,
                                    self.loss.kind, margin, adv, temp, kw.get("p_norm", 1), kw.get("normalize", False))
