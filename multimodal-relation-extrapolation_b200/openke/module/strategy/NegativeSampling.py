"""NegativeSampling strategy (OpenKE/openke/module/strategy/NegativeSampling.py:5-32): splits the flat score vector
into p_score [B, 1] and n_score [B, neg] and applies the loss.  For TransE + plain MarginLoss with no regulariser
(the configuration of OpenKE/examples/train_transe_FB15K237.py:23-39, BASELINE configs[3]) `fused_step` runs the whole
forward + backward in mre_transe_margin_step and leaves the gradients in .grad, so Trainer skips autograd."""
import torch
import torch.nn as nn

from .... import engine
from ..loss.MarginLoss import MarginLoss
from ..model.Model import expand_batch
from ..model.TransE import TransE


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
        return (isinstance(self.model, TransE) and isinstance(self.loss, MarginLoss) and not self.loss.adv_flag
                and not self.model.margin_flag and self.regul_rate == 0 and self.l3_regul_rate == 0)

    def fused_step(self, data):
        """loss [1] of one batch with d loss / d tables ACCUMULATED into the embeddings' .grad (as backward() would)"""
        m = self.model
        ent, rel = m.tables()
        h, t, r = expand_batch(data, ent.device)
        n = h.numel()
        B = self.batch_size
        assert n % B == 0 and n > B
        for w in (ent, rel):
            if w.grad is None:
                w.grad = torch.zeros_like(w)
        loss, _, _, _ = engine.transe_margin_step(m.ctx(), ent.data, rel.data, h, t, r, B, n // B - 1, float(self.loss.margin.item()),
                                                  m.p_norm, m.norm_flag, grad_ent=ent.grad, grad_rel=rel.grad)
        return loss
