"""TransE (OpenKE/openke/module/model/TransE.py): same constructor, same parameter names and initialisation, same
forward / predict / regularization semantics; scoring runs in libmre_b200.so."""
import torch
import torch.nn as nn

from .Model import Model


class TransE(Model):
    scorer = "transe"

    def __init__(self, ent_tot, rel_tot, dim=100, p_norm=1, norm_flag=True, margin=None, epsilon=None):
        super().__init__(ent_tot, rel_tot)
        self.dim = dim
        self.margin = margin
        self.epsilon = epsilon
        self.norm_flag = norm_flag
        self.p_norm = p_norm
        self.ent_embeddings = nn.Embedding(self.ent_tot, self.dim)
        self.rel_embeddings = nn.Embedding(self.rel_tot, self.dim)
        if margin is None or epsilon is None:                                   # TransE.py:20-22
            nn.init.xavier_uniform_(self.ent_embeddings.weight.data)
            nn.init.xavier_uniform_(self.rel_embeddings.weight.data)
        else:                                                                   # TransE.py:23-38
            self.embedding_range = nn.Parameter(torch.Tensor([(self.margin + self.epsilon) / self.dim]), requires_grad=False)
            nn.init.uniform_(tensor=self.ent_embeddings.weight.data, a=-self.embedding_range.item(), b=self.embedding_range.item())
            nn.init.uniform_(tensor=self.rel_embeddings.weight.data, a=-self.embedding_range.item(), b=self.embedding_range.item())
        if margin is not None:                                                  # TransE.py:40-44
            self.margin = nn.Parameter(torch.Tensor([margin]))
            self.margin.requires_grad = False
            self.margin_flag = True
        else:
            self.margin_flag = False

    def tables(self):
        return self.ent_embeddings.weight, self.rel_embeddings.weight

    def rank_kwargs(self):
        return {"p_norm": self.p_norm, "normalize": self.norm_flag}

    def forward(self, data):                                                    # TransE.py:62-74
        score = self.raw_score(data)
        if self.margin_flag:
            return self.margin - score
        return score

    def regularization(self, data):                                             # TransE.py:76-86
        h = self.ent_embeddings(data["batch_h"])
        t = self.ent_embeddings(data["batch_t"])
        r = self.rel_embeddings(data["batch_r"])
        return (torch.mean(h ** 2) + torch.mean(t ** 2) + torch.mean(r ** 2)) / 3

    def predict(self, data):                                                    # TransE.py:88-94: the distance, lower is better
        with torch.no_grad():
            score = self.raw_score(data)
        return score.cpu().data.numpy()
