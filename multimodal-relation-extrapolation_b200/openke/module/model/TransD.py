"""TransD (OpenKE/openke/module/model/TransD.py): entities move by a rank-one map built from an entity and a relation transfer
vector, then are normalised, before the translation.  Same constructor, parameter names (ent_embeddings, rel_embeddings,
ent_transfer, rel_transfer) and initialisation; forward / predict follow :111-155.  Link-prediction ranking runs on the TransE
kernel over per-relation projected tables (_projected.py; dim_e == dim_r); forward() on explicit triples is the same arithmetic
in torch."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .... import _lib as L
from .Model import Model, expand_batch
from ._projected import RelationProjected


class TransD(RelationProjected, Model):
    scorer = None
    project_kind = L.PROJECT_TRANSD

    def __init__(self, ent_tot, rel_tot, dim_e=100, dim_r=100, p_norm=1, norm_flag=True, margin=None, epsilon=None):
        super().__init__(ent_tot, rel_tot)
        self.dim_e, self.dim_r, self.margin, self.epsilon, self.norm_flag, self.p_norm = dim_e, dim_r, margin, epsilon, norm_flag, p_norm
        self.ent_embeddings = nn.Embedding(ent_tot, dim_e)
        self.rel_embeddings = nn.Embedding(rel_tot, dim_r)
        self.ent_transfer = nn.Embedding(ent_tot, dim_e)
        self.rel_transfer = nn.Embedding(rel_tot, dim_r)
        tabs = (self.ent_embeddings, self.rel_embeddings, self.ent_transfer, self.rel_transfer)
        if margin is None or epsilon is None:                                   # TransD.py:22-26
            for e in tabs:
                nn.init.xavier_uniform_(e.weight.data)
        else:                                                                   # TransD.py:27-54
            self.ent_embedding_range = nn.Parameter(torch.Tensor([(margin + epsilon) / dim_e]), requires_grad=False)
            self.rel_embedding_range = nn.Parameter(torch.Tensor([(margin + epsilon) / dim_r]), requires_grad=False)
            for e, rng in zip(tabs, (self.ent_embedding_range, self.rel_embedding_range) * 2):
                nn.init.uniform_(tensor=e.weight.data, a=-rng.item(), b=rng.item())
        self.margin_flag = margin is not None
        if self.margin_flag:
            self.margin = nn.Parameter(torch.Tensor([margin]), requires_grad=False)

    def projection_tables(self):
        return self.ent_embeddings.weight, self.ent_transfer.weight, self.rel_embeddings.weight, self.rel_transfer.weight

    def _fit(self, x):
        """_resize (:62-75): cut or zero-pad the entity vector to the relation space"""
        if self.dim_e == self.dim_r:
            return x
        return x[..., :self.dim_r] if self.dim_e > self.dim_r else F.pad(x, (0, self.dim_r - self.dim_e))

    def _move(self, e, e_p, r_p):                                               # _transfer, :92-109
        return F.normalize(self._fit(e) + (e * e_p).sum(-1, keepdim=True) * r_p, p=2, dim=-1)

    def forward(self, data):
        h, t, r = expand_batch(data, self.device())
        r_p = self.rel_transfer(r)
        hh = self._move(self.ent_embeddings(h), self.ent_transfer(h), r_p)
        tt = self._move(self.ent_embeddings(t), self.ent_transfer(t), r_p)
        rr = self.rel_embeddings(r)
        if self.norm_flag:                                                      # _calc, :77-90
            hh, rr, tt = F.normalize(hh, 2, -1), F.normalize(rr, 2, -1), F.normalize(tt, 2, -1)
        u = hh + (rr - tt) if data["mode"] == "head_batch" else (hh + rr) - tt
        score = torch.norm(u, self.p_norm, -1)
        return self.margin - score if self.margin_flag else score

    def regularization(self, data):                                             # TransD.py:133-147
        parts = (self.ent_embeddings(data["batch_h"]), self.ent_embeddings(data["batch_t"]), self.rel_embeddings(data["batch_r"]),
                 self.ent_transfer(data["batch_h"]), self.ent_transfer(data["batch_t"]), self.rel_transfer(data["batch_r"]))
        return sum(torch.mean(x ** 2) for x in parts) / 6

    def predict(self, data):                                                    # TransD.py:149-155
        with torch.no_grad():
            score = self.forward(data)
            if self.margin_flag:
                score = self.margin - score
        return score.cpu().data.numpy()
