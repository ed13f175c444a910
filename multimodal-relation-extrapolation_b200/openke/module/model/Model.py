"""Base class of the mirrored scorers (OpenKE/openke/module/model/Model.py:5-16, BaseModule.py:16-55): an nn.Module
holding the same embedding tables under the same names, whose forward/predict call the library's kernels."""
import json
import os

import numpy as np
import torch
import torch.nn as nn

from .... import engine
from .... import _lib as L


def _as_index(x, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.to(device=device, dtype=torch.int64).contiguous()


def expand_batch(data, device):
    """index arrays of one loader batch expanded to explicit triples of equal length n (TransE.py:51-54's
    view(-1, r.shape[0], D) broadcast for head_batch / tail_batch), on `device`"""
    h, t, r = (_as_index(data[k], device) for k in ("batch_h", "batch_t", "batch_r"))
    n = max(h.numel(), t.numel(), r.numel())

    def grow(x):
        if x.numel() == n:
            return x
        assert n % x.numel() == 0
        return x.repeat(n // x.numel())
    return grow(h), grow(t), grow(r)


def _check_ids(h, t, r, E, R):
    """torch's embedding lookup raises on an out-of-range id; the kernels would read (and atomically add) out of bounds"""
    if h.numel() == 0:
        return
    bad = ((h < 0) | (h >= E)).any() | ((t < 0) | (t >= E)).any() | ((r < 0) | (r >= R)).any()
    if bool(bad):                                                               # one small device read per call
        raise IndexError(f"triple ids out of range for {E} entities / {R} relations")


class _ScoreFn(torch.autograd.Function):
    """raw model score of explicit triples with gradients to every embedding table of the model: mre_score_triples forward,
    mre_transe_backward (TransE) / mre_bilinear_backward (DistMult, SimplE, ComplEx) backward"""

    @staticmethod
    def forward(ctx, model, h, t, r, *tabs):
        score = model._score(h, t, r, tabs)
        ctx.model, ctx.idx = model, (h, t, r)
        ctx.save_for_backward(score, *tabs)
        return score

    @staticmethod
    def backward(ctx, dscore):
        score, *tabs = ctx.saved_tensors
        m = ctx.model
        h, t, r = ctx.idx
        lib, st = L.lib(), torch.cuda.current_stream().cuda_stream
        grads = [torch.zeros_like(w) for w in tabs]
        dscore = dscore.contiguous()
        D = tabs[0].shape[1]
        if m.scorer == "transe":
            ent, rel = tabs
            L.check(lib.mre_transe_backward(m.ctx()._h, ent.data_ptr(), rel.data_ptr(), D, h.data_ptr(), t.data_ptr(), r.data_ptr(),
                                            h.numel(), m.p_norm, int(m.norm_flag), score.data_ptr(), dscore.data_ptr(),
                                            grads[0].data_ptr(), grads[1].data_ptr(), st))
        else:
            if m.scorer == "complex":
                ent, ent_im, rel, rel_im = tabs
                g_ent, g_ent_im, g_rel, g_rel_im = grads
            else:
                (ent, rel), ent_im, rel_im = tabs, None, None
                (g_ent, g_rel), g_ent_im, g_rel_im = grads, None, None
            ptr = lambda x: x.data_ptr() if x is not None else None
            L.check(lib.mre_bilinear_backward(m.ctx()._h, engine.SCORERS[m.scorer], ent.data_ptr(), ptr(ent_im), rel.data_ptr(), ptr(rel_im),
                                              D, h.data_ptr(), t.data_ptr(), r.data_ptr(), h.numel(), dscore.data_ptr(),
                                              g_ent.data_ptr(), ptr(g_ent_im), g_rel.data_ptr(), ptr(g_rel_im), st))
        return (None, None, None, None, *grads)


class Model(nn.Module):
    scorer = None

    def __init__(self, ent_tot, rel_tot):
        super().__init__()
        self.ent_tot = ent_tot
        self.rel_tot = rel_tot
        self.zero_const = nn.Parameter(torch.Tensor([0]), requires_grad=False)
        self.pi_const = nn.Parameter(torch.Tensor([3.14159265358979323846]), requires_grad=False)
        self._ctx = None
        self._ranker = None

    # ---- library plumbing
    def device(self):
        return next(self.parameters()).device

    def ctx(self):
        dev = self.device()
        if dev.type != "cuda":
            raise L.MreError("mre_b200 models compute on a B200 only: call model.cuda() first (there is no CPU fallback)")
        if self._ctx is None or self._ctx.device != (dev.index or 0):
            self._ctx = engine.Context(dev.index or 0)
            self._ranker = engine.Ranker(self._ctx)
        return self._ctx

    def ranker(self):
        self.ctx()
        return self._ranker

    def tables(self):
        raise NotImplementedError

    def rank_kwargs(self):
        return {}

    def _score(self, h, t, r, tabs=None):
        tabs = tabs if tabs is not None else self.tables()
        if self.scorer == "complex":
            ent, ent_im, rel, rel_im = tabs
        else:
            (ent, rel), ent_im, rel_im = tabs, None, None
        out = torch.empty(h.numel(), dtype=torch.float32, device=ent.device)
        kw = self.rank_kwargs()
        L.check(L.lib().mre_score_triples(self.ctx()._h, engine.SCORERS[self.scorer], ent.data_ptr(),
                                          ent_im.data_ptr() if ent_im is not None else None, rel.data_ptr(),
                                          rel_im.data_ptr() if rel_im is not None else None, ent.shape[1], h.data_ptr(), t.data_ptr(),
                                          r.data_ptr(), h.numel(), kw.get("p_norm", 1), int(kw.get("normalize", False)), out.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream))
        return out

    def raw_score(self, data, tables=None):
        """the model's raw score (distance for TransE, similarity for DistMult/ComplEx) of the batch's triples, differentiable
        with respect to the embedding tables (`tables` overrides the model's own, e.g. SimplE's inverse-relation pass)"""
        h, t, r = expand_batch(data, self.device())
        tabs = tuple(tables) if tables is not None else tuple(self.tables())
        _check_ids(h, t, r, tabs[0].shape[0], tabs[-1].shape[0])
        if torch.is_grad_enabled() and any(p.requires_grad for p in tabs):
            return _ScoreFn.apply(self, h, t, r, *tabs)
        return self._score(h, t, r, tabs)

    # ---- BaseModule.py:16-55
    def load_checkpoint(self, path):
        self.load_state_dict(torch.load(os.path.join(path)))
        self.eval()

    def save_checkpoint(self, path):
        torch.save(self.state_dict(), path)

    def load_parameters(self, path):
        with open(path, "r") as f:
            parameters = json.loads(f.read())
        for i in parameters:
            parameters[i] = torch.Tensor(parameters[i])
        self.load_state_dict(parameters, strict=False)
        self.eval()

    def save_parameters(self, path):
        with open(path, "w") as f:
            f.write(json.dumps(self.get_parameters("list")))

    def get_parameters(self, mode="numpy", param_dict=None):
        all_param_dict = self.state_dict()
        if param_dict is None:
            param_dict = all_param_dict.keys()
        res = {}
        for param in param_dict:
            if mode == "numpy":
                res[param] = all_param_dict[param].cpu().numpy()
            elif mode == "list":
                res[param] = all_param_dict[param].cpu().numpy().tolist()
            else:
                res[param] = all_param_dict[param]
        return res

    def set_parameters(self, parameters):
        for i in parameters:
            parameters[i] = torch.Tensor(parameters[i])
        self.load_state_dict(parameters, strict=False)
        self.eval()
