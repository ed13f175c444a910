"""torch-CPU restatement of the reference's Python scoring/loss code -- TEST INFRASTRUCTURE ONLY.

The reference's own modules (OpenKE/openke/module/...) can be imported in the build container but do
not travel to the GPU box, so the same tensor expressions are restated here.  tests/golden/make_golden.py
asserts, in the build container, that every function below returns BIT-IDENTICAL tensors to the real
reference module on the same inputs (same torch build), which pins this file to the reference.
It is also the "plain PyTorch fp32 reference" for the floating-point training kernel.
"""
import torch
import torch.nn.functional as F


def transe_calc(h, t, r, mode, p_norm=1, norm_flag=True):
    """OpenKE/openke/module/model/TransE.py:46-60 (also module/NegativeSampling.py:142-157)."""
    if norm_flag:
        h = F.normalize(h, 2, -1)
        r = F.normalize(r, 2, -1)
        t = F.normalize(t, 2, -1)
    if mode != "normal":
        h = h.view(-1, r.shape[0], h.shape[-1])
        t = t.view(-1, r.shape[0], t.shape[-1])
        r = r.view(-1, r.shape[0], r.shape[-1])
    if mode == "head_batch":
        score = h + (r - t)
    else:
        score = (h + r) - t
    return torch.norm(score, p_norm, -1).flatten()


def distmult_calc(h, t, r, mode):
    """OpenKE/openke/module/model/DistMult.py:34-44 (also module/NegativeSampling.py:158-168)."""
    if mode != "normal":
        h = h.view(-1, r.shape[0], h.shape[-1])
        t = t.view(-1, r.shape[0], t.shape[-1])
        r = r.view(-1, r.shape[0], r.shape[-1])
    if mode == "head_batch":
        score = h * (r * t)
    else:
        score = (h * r) * t
    return torch.sum(score, -1).flatten()


def complex_calc(h_re, h_im, t_re, t_im, r_re, r_im):
    """OpenKE/openke/module/model/ComplEx.py:20-27."""
    return torch.sum(h_re * t_re * r_re + h_im * t_im * r_re + h_re * t_im * r_im - h_im * t_re * r_im, -1)


def predict(kind, tables, data, p_norm=1, norm_flag=True):
    """`Model.predict` (TransE.py:88-94, DistMult.py:70-72, ComplEx.py:60-61): lower is better.

    tables: TransE/DistMult (ent, rel); ComplEx (ent_re, ent_im, rel_re, rel_im).  data: the dict the
    loaders yield ({batch_h, batch_t, batch_r, mode}) with torch int64 tensors.
    """
    bh, bt, br, mode = data["batch_h"], data["batch_t"], data["batch_r"], data["mode"]
    if kind == "transe":
        ent, rel = tables
        return transe_calc(ent[bh], ent[bt], rel[br], mode, p_norm, norm_flag)
    if kind == "distmult":
        ent, rel = tables
        return -distmult_calc(ent[bh], ent[bt], rel[br], mode)
    if kind == "complex":
        ent_re, ent_im, rel_re, rel_im = tables
        return -complex_calc(ent_re[bh], ent_im[bh], ent_re[bt], ent_im[bt], rel_re[br], rel_im[br])
    raise ValueError(kind)


def margin_loss(p_score, n_score, margin):
    """OpenKE/openke/module/loss/MarginLoss.py:24-28 (no adversarial weights); copy at module/loss.py:20-24."""
    m = torch.tensor([margin], dtype=p_score.dtype)
    return (torch.max(p_score - n_score, -m)).mean() + m


def strategy_loss(score, batch_size, margin):
    """OpenKE/openke/module/strategy/NegativeSampling.py:13-32 with regul_rate = l3_regul_rate = 0."""
    p = score[:batch_size].view(-1, batch_size).permute(1, 0)
    n = score[batch_size:].view(-1, batch_size).permute(1, 0)
    return margin_loss(p, n, margin)


def transe_train_step(ent, rel, bh, bt, br, batch_size, margin, p_norm=1, norm_flag=True):
    """Loss and dense gradients of one OpenKE TransE step (Trainer.py:43-54 without the optimiser)."""
    ent = ent.clone().requires_grad_(True)
    rel = rel.clone().requires_grad_(True)
    score = transe_calc(ent[bh], ent[bt], rel[br], "normal", p_norm, norm_flag)
    loss = strategy_loss(score, batch_size, margin)
    loss.backward()
    return loss.detach(), score.detach(), ent.grad, rel.grad
