"""ComplEx (OpenKE/openke/module/model/ComplEx.py)."""
import torch
import torch.nn as nn

from .Model import Model


class ComplEx(Model):
    scorer = "complex"

    def __init__(self, ent_tot, rel_tot, dim=100):
        super().__init__(ent_tot, rel_tot)
        self.dim = dim
        self.ent_re_embeddings = nn.Embedding(self.ent_tot, self.dim)
        self.ent_im_embeddings = nn.Embedding(self.ent_tot, self.dim)
        self.rel_re_embeddings = nn.Embedding(self.rel_tot, self.dim)
        self.rel_im_embeddings = nn.Embedding(self.rel_tot, self.dim)
        for e in (self.ent_re_embeddings, self.ent_im_embeddings, self.rel_re_embeddings, self.rel_im_embeddings):
            nn.init.xavier_uniform_(e.weight.data)                              # ComplEx.py:15-18

    def tables(self):
        return (self.ent_re_embeddings.weight, self.ent_im_embeddings.weight, self.rel_re_embeddings.weight,
                self.rel_im_embeddings.weight)

    def forward(self, data):                                                    # ComplEx.py:29-40
        return self.raw_score(data)

    def regularization(self, data):                                             # ComplEx.py:42-57
        bh, bt, br = data["batch_h"], data["batch_t"], data["batch_r"]
        parts = [self.ent_re_embeddings(bh), self.ent_im_embeddings(bh), self.ent_re_embeddings(bt), self.ent_im_embeddings(bt),
                 self.rel_re_embeddings(br), self.rel_im_embeddings(br)]
        return sum(torch.mean(p ** 2) for p in parts) / 6

    def predict(self, data):                                                    # ComplEx.py:60-61
        with torch.no_grad():
            score = -self.raw_score(data)
        return score.cpu().data.numpy()
