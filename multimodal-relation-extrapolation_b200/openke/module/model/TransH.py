"""TransH (OpenKE/openke/module/model/TransH.py): entities are projected onto the relation's hyperplane before the translation.
Same constructor, parameter names (ent_embeddings, rel_embeddings, norm_vector) and initialisation; forward / predict follow
:76-117.  Link-prediction ranking runs on the TransE kernel over per-relation projected tables (_projected.py); forward() on
explicit triples is the same arithmetic in torch (differentiable: a model trains through it with any of the library's losses)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .... import _lib as L
from .Model import Model, expand_batch
from ._projected import RelationProjected


class TransH(RelationProjected, Model):
    scorer = None                      # no single-table scorer: ranks through rank_queries()
    project_kind = L.PROJECT_TRANSH

    def __init__(self, ent_tot, rel_tot, dim=100, p_norm=1, norm_flag=True, margin=None, epsilon=None):
        super().__init__(ent_tot, rel_tot)
        self.dim, self.margin, self.epsilon, self.norm_flag, self.p_norm = dim, margin, epsilon, norm_flag, p_norm
        self.ent_embeddings = nn.Embedding(ent_tot, dim)
        self.rel_embeddings = nn.Embedding(rel_tot, dim)
        self.norm_vector = nn.Embedding(rel_tot, dim)
        tabs = (self.ent_embeddings, self.rel_embeddings, self.norm_vector)
        if margin is None or epsilon is None:                                   # TransH.py:21-24
            for e in tabs:
                nn.init.xavier_uniform_(e.weight.data)
        else:                                                                   # TransH.py:25-43
            self.embedding_range = nn.Parameter(torch.Tensor([(margin + epsilon) / dim]), requires_grad=False)
            for e in tabs:
                nn.init.uniform_(tensor=e.weight.data, a=-self.embedding_range.item(), b=self.embedding_range.item())
        self.margin_flag = margin is not None                                   # TransH.py:45-50
        if self.margin_flag:
            self.margin = nn.Parameter(torch.Tensor([margin]), requires_grad=False)

    def projection_tables(self):
        return self.ent_embeddings.weight, None, self.rel_embeddings.weight, self.norm_vector.weight

    def _distance(self, h, t, r, w, mode):
        w = F.normalize(w, p=2, dim=-1)
        h = h - (h * w).sum(-1, keepdim=True) * w                               # _transfer, :66-74
        t = t - (t * w).sum(-1, keepdim=True) * w
        if self.norm_flag:                                                      # _calc, :51-64
            h, r, t = F.normalize(h, 2, -1), F.normalize(r, 2, -1), F.normalize(t, 2, -1)
        u = h + (r - t) if mode == "head_batch" else (h + r) - t
        return torch.norm(u, self.p_norm, -1)

    def forward(self, data):
        h, t, r = expand_batch(data, self.device())
        score = self._distance(self.ent_embeddings(h), self.ent_embeddings(t), self.rel_embeddings(r), self.norm_vector(r), data["mode"])
        return self.margin - score if self.margin_flag else score

    def regularization(self, data):                                             # TransH.py:98-107
        parts = (self.ent_embeddings(data["batch_h"]), self.ent_embeddings(data["batch_t"]), self.rel_embeddings(data["batch_r"]),
                 self.norm_vector(data["batch_r"]))
        return sum(torch.mean(x ** 2) for x in parts) / 4

    def predict(self, data):                                                    # TransH.py:109-115: the distance, lower is better
        with torch.no_grad():
            score = self.forward(data)
            if self.margin_flag:
                score = self.margin - score
        return score.cpu().data.numpy()
