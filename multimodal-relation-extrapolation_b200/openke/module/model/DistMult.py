"""DistMult (OpenKE/openke/module/model/DistMult.py)."""
import torch
import torch.nn as nn

from .Model import Model


class DistMult(Model):
    scorer = "distmult"

    def __init__(self, ent_tot, rel_tot, dim=100, margin=None, epsilon=None):
        super().__init__(ent_tot, rel_tot)
        self.dim = dim
        self.margin = margin
        self.epsilon = epsilon
        self.ent_embeddings = nn.Embedding(self.ent_tot, self.dim)
        self.rel_embeddings = nn.Embedding(self.rel_tot, self.dim)
        if margin is None or epsilon is None:                                   # DistMult.py:17-19
            nn.init.xavier_uniform_(self.ent_embeddings.weight.data)
            nn.init.xavier_uniform_(self.rel_embeddings.weight.data)
        else:                                                                   # DistMult.py:20-32
            self.embedding_range = nn.Parameter(torch.Tensor([(self.margin + self.epsilon) / self.dim]), requires_grad=False)
            nn.init.uniform_(tensor=self.ent_embeddings.weight.data, a=-self.embedding_range.item(), b=self.embedding_range.item())
            nn.init.uniform_(tensor=self.rel_embeddings.weight.data, a=-self.embedding_range.item(), b=self.embedding_range.item())

    def tables(self):
        return self.ent_embeddings.weight, self.rel_embeddings.weight

    def forward(self, data):                                                    # DistMult.py:46-57: the similarity
        return self.raw_score(data)

    def regularization(self, data):                                             # DistMult.py:59-65
        h = self.ent_embeddings(data["batch_h"])
        t = self.ent_embeddings(data["batch_t"])
        r = self.rel_embeddings(data["batch_r"])
        return (torch.mean(h ** 2) + torch.mean(t ** 2) + torch.mean(r ** 2)) / 3

    def l3_regularization(self):                                                # DistMult.py:67-68
        return self.ent_embeddings.weight.norm(p=3) ** 3 + self.rel_embeddings.weight.norm(p=3) ** 3

    def predict(self, data):                                                    # DistMult.py:70-72: -score, lower is better
        with torch.no_grad():
            score = -self.raw_score(data)
        return score.cpu().data.numpy()
