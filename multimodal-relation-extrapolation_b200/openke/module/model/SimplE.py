"""SimplE (OpenKE/openke/module/model/SimplE.py).  Its link-prediction scorer is the trilinear product of the FORWARD tables
only -- predict = -sum(h * r * t) (:48-55; rel_inv_embeddings enter the training score _calc_avg :19-20, not predict) -- so
ranking runs on the same tcgen05 contraction as DistMult (scorer "distmult" over ent_embeddings / rel_embeddings)."""
import torch
import torch.nn as nn

from .Model import Model


class SimplE(Model):
    scorer = "distmult"

    def __init__(self, ent_tot, rel_tot, dim=100):
        super().__init__(ent_tot, rel_tot)
        self.dim = dim
        self.ent_embeddings = nn.Embedding(self.ent_tot, self.dim)
        self.rel_embeddings = nn.Embedding(self.rel_tot, self.dim)
        self.rel_inv_embeddings = nn.Embedding(self.rel_tot, self.dim)
        nn.init.xavier_uniform_(self.ent_embeddings.weight.data)
        nn.init.xavier_uniform_(self.rel_embeddings.weight.data)
        nn.init.xavier_uniform_(self.rel_inv_embeddings.weight.data)

    def tables(self):
        return self.ent_embeddings.weight, self.rel_embeddings.weight

    def forward(self, data):                                                    # SimplE.py:25-34: (<h,r,t> + <h,r_inv,t>) / 2
        fwd = self.raw_score(data)
        inv = self.raw_score(data, tables=(self.ent_embeddings.weight, self.rel_inv_embeddings.weight))
        return (fwd + inv) / 2

    def regularization(self, data):                                             # SimplE.py:36-45
        h = self.ent_embeddings(data["batch_h"])
        t = self.ent_embeddings(data["batch_t"])
        r = self.rel_embeddings(data["batch_r"])
        r_inv = self.rel_inv_embeddings(data["batch_r"])
        return (torch.mean(h ** 2) + torch.mean(t ** 2) + torch.mean(r ** 2) + torch.mean(r_inv ** 2)) / 4

    def predict(self, data):                                                    # SimplE.py:47-55: -<h,r,t>, lower is better
        with torch.no_grad():
            score = -self.raw_score(data)
        return score.cpu().data.numpy()
