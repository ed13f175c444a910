"""RotatE (OpenKE/openke/module/model/RotatE.py): entities are complex vectors [re | im] of width 2 dim, a relation is a vector
of dim phases; the score of a triple is the sum over dimensions of | h o r - t | (complex modulus).  Same constructor, parameter
names (ent_embeddings, rel_embeddings, ent_embedding_range, rel_embedding_range, margin) and initialisation (:8-42).
Link-prediction ranking runs in the library's RotatE tile kernel (csrc/rotate_rank.cu, scorer "rotate"); forward() on explicit
triples is the same arithmetic in torch (differentiable: the model trains through it with any of the library's losses)."""
import torch
import torch.nn as nn

from .Model import Model, expand_batch


class RotatE(Model):
    scorer = "rotate"
    fusable = False                    # no fused training step for this scorer: the strategy runs forward() + autograd

    def __init__(self, ent_tot, rel_tot, dim=100, margin=6.0, epsilon=2.0):
        super().__init__(ent_tot, rel_tot)
        self.epsilon = epsilon
        self.dim_e, self.dim_r = dim * 2, dim
        self.ent_embeddings = nn.Embedding(ent_tot, self.dim_e)
        self.rel_embeddings = nn.Embedding(rel_tot, self.dim_r)
        self.ent_embedding_range = nn.Parameter(torch.Tensor([(margin + epsilon) / self.dim_e]), requires_grad=False)
        self.rel_embedding_range = nn.Parameter(torch.Tensor([(margin + epsilon) / self.dim_r]), requires_grad=False)
        nn.init.uniform_(tensor=self.ent_embeddings.weight.data, a=-self.ent_embedding_range.item(), b=self.ent_embedding_range.item())
        nn.init.uniform_(tensor=self.rel_embeddings.weight.data, a=-self.rel_embedding_range.item(), b=self.rel_embedding_range.item())
        self.margin = nn.Parameter(torch.Tensor([margin]), requires_grad=False)

    def tables(self):
        return self.ent_embeddings.weight, self.rel_embeddings.weight

    def phase_div(self):
        """rel_embedding_range / pi in float32, the divisor of RotatE.py:49 (.item() is a double, pi_const a float32 tensor)"""
        return float((self.rel_embedding_range.item() / self.pi_const.detach().cpu()).item())

    def rank_kwargs(self):
        return {"phase_div": self.phase_div()}

    def _distance(self, h, t, r, mode):                                         # RotatE.py:44-78 on explicit triples
        re_h, im_h = torch.chunk(h, 2, dim=-1)
        re_t, im_t = torch.chunk(t, 2, dim=-1)
        phase = r / (self.rel_embedding_range.item() / self.pi_const)
        re_r, im_r = torch.cos(phase), torch.sin(phase)
        if mode == "head_batch":                                                # conj(r) o t - h
            re_s = (re_r * re_t + im_r * im_t) - re_h
            im_s = (re_r * im_t - im_r * re_t) - im_h
        else:                                                                   # h o r - t
            re_s = (re_h * re_r - im_h * im_r) - re_t
            im_s = (re_h * im_r + im_h * re_r) - im_t
        return torch.stack([re_s, im_s], dim=0).norm(dim=0).sum(dim=-1)

    def forward(self, data):                                                    # RotatE.py:80-88
        h, t, r = expand_batch(data, self.device())
        return self.margin - self._distance(self.ent_embeddings(h), self.ent_embeddings(t), self.rel_embeddings(r), data["mode"])

    def regularization(self, data):                                             # RotatE.py:94-103
        h = self.ent_embeddings(data["batch_h"])
        t = self.ent_embeddings(data["batch_t"])
        r = self.rel_embeddings(data["batch_r"])
        return (torch.mean(h ** 2) + torch.mean(t ** 2) + torch.mean(r ** 2)) / 3

    def predict(self, data):                                                    # RotatE.py:90-92: distance - margin, lower is better
        with torch.no_grad():
            score = -self.forward(data)
        return score.cpu().data.numpy()
