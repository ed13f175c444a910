"""Analogy (OpenKE/openke/module/model/Analogy.py): a ComplEx block of width dim plus a DistMult block of width 2 dim.
Same constructor, parameter names and initialisation.  The score is one trilinear form, so it runs on the ComplEx skeleton
(tcgen05 contraction, K = 2 * 3 dim) over widened tables: real parts [e_re | e], imaginary parts [e_im | 0] -- the appended
dimensions have no imaginary part, where ComplEx's four-term product collapses to h * r * t.  forward (:29-45) returns MINUS the
similarity and predict (:74-76) returns minus forward, i.e. the similarity itself with lower = better (the model is trained to
make forward() large for true triples); the library's ComplEx predict is minus the similarity, so the relation tables enter negated."""
import torch
import torch.nn as nn

from .Model import Model


class Analogy(Model):
    scorer = "complex"
    fusable = False                    # tables() are derived tensors: the fused step's in-place gradients would miss the parameters

    def __init__(self, ent_tot, rel_tot, dim=100):
        super().__init__(ent_tot, rel_tot)
        self.dim = dim
        self.ent_re_embeddings = nn.Embedding(ent_tot, dim)
        self.ent_im_embeddings = nn.Embedding(ent_tot, dim)
        self.rel_re_embeddings = nn.Embedding(rel_tot, dim)
        self.rel_im_embeddings = nn.Embedding(rel_tot, dim)
        self.ent_embeddings = nn.Embedding(ent_tot, dim * 2)
        self.rel_embeddings = nn.Embedding(rel_tot, dim * 2)
        for e in (self.ent_re_embeddings, self.ent_im_embeddings, self.rel_re_embeddings, self.rel_im_embeddings, self.ent_embeddings,
                  self.rel_embeddings):
            nn.init.xavier_uniform_(e.weight.data)

    def tables(self):
        """(ent_re', ent_im', rel_re', rel_im') of the equivalent ComplEx model of width 3 dim whose predict() is Analogy's"""
        z_e = torch.zeros_like(self.ent_embeddings.weight)
        z_r = torch.zeros_like(self.rel_embeddings.weight)
        return (torch.cat([self.ent_re_embeddings.weight, self.ent_embeddings.weight], 1), torch.cat([self.ent_im_embeddings.weight, z_e], 1),
                -torch.cat([self.rel_re_embeddings.weight, self.rel_embeddings.weight], 1), -torch.cat([self.rel_im_embeddings.weight, z_r], 1))

    def forward(self, data):                                                    # Analogy.py:29-45: -(complex similarity) - (distmult similarity)
        # raw_score over the widened tables is the similarity with the relation negated = forward()'s value; differentiable through
        # torch.cat back to the six embedding tables
        return self.raw_score(data, tables=self.tables())

    def regularization(self, data):                                             # Analogy.py:47-71
        h, t, r = data["batch_h"], data["batch_t"], data["batch_r"]
        parts = (self.ent_re_embeddings(h), self.ent_im_embeddings(h), self.ent_embeddings(h), self.ent_re_embeddings(t), self.ent_im_embeddings(t),
                 self.ent_embeddings(t), self.rel_re_embeddings(r), self.rel_im_embeddings(r), self.rel_embeddings(r))
        return sum(torch.mean(x ** 2) for x in parts) / 9

    def predict(self, data):                                                    # Analogy.py:73-75
        with torch.no_grad():
            score = -self.forward(data)
        return score.cpu().data.numpy()
