"""Restatement of the ZSL candidate scorer -- TEST INFRASTRUCTURE ONLY.

What it restates (paths relative to /root/reference):
  module/zsl_module.py:46-59    Extractor.neighbor_encoder: gcn_w over the 50 neighbour symbols, sum / degree, tanh
  module/zsl_module.py:61-67    Extractor.entity_encoder: tanh([fc1(e1) | fc2(e2)])
  module/zsl_module.py:69-106   Extractor.forward (query half): [left | pair | right] -> reshape_layer -> SupportEncoder
  module/submodule.py:240-258   SupportEncoder: LayerNorm(proj2(relu(proj1(x))) + x)   (eval mode: dropout off)
  module/zsl_module.py:662-706  ZSLmodule.eval: candidate_vecs vs the 20 generated relation vectors, sklearn
                                cosine_similarity(...).mean(axis=1), rank = 1 + position of candidate 0 in the descending argsort
Pinning: tests/golden/make_golden_zsl.py imports the reference's own Extractor (module.zsl_module, with the three modules it
cannot import here stubbed) on seeded weights and asserts `extractor_query_vectors` below is bit-identical to it; the fixture
tests/golden/golden_zsl.npz holds the reference's scores and ranks.  `separable_scores` is the algebraically equal form the
CUDA path computes (per-entity halves A_h + B_c); it differs from the reference by FP32 rounding only.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

D_DEFAULT = 200

# the seeded input generators live with the fixtures (tests/golden/golden_util.py: shared by tests and bench.py, which may not
# import this package outside its cpu_baseline leg); re-exported here for the golden scripts
_gdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
if _gdir not in sys.path:
    sys.path.insert(0, _gdir)
from golden_util import seeded_extractor_weights, synthetic_zsl_setup  # noqa: E402,F401


def extractor_query_vectors(w, pairs, left_conn, left_deg, right_conn, right_deg):
    """Extractor.forward's query_g for (head symbol, tail symbol) pairs -- the same torch expressions, eval mode."""
    t = {k: torch.from_numpy(v) for k, v in w.items()}
    emb = t["symbol_emb.weight"]
    pairs = torch.as_tensor(pairs, dtype=torch.long)
    def neighbor(conn, deg):
        ent = emb[torch.as_tensor(conn, dtype=torch.long)]                       # (batch, max_nb, D)
        out = F.linear(ent, t["gcn_w.weight"], t["gcn_w.bias"])
        out = torch.sum(out, dim=1)
        out = out / torch.as_tensor(deg, dtype=torch.float32).unsqueeze(1)
        return out.tanh()
    e1, e2 = emb[pairs[:, 0]], emb[pairs[:, 1]]
    ent = torch.cat((F.linear(e1, t["fc1.weight"], t["fc1.bias"]), F.linear(e2, t["fc2.weight"], t["fc2.bias"])), dim=-1).tanh()
    q = torch.cat((neighbor(left_conn, left_deg), ent, neighbor(right_conn, right_deg)), dim=-1)
    q = F.linear(q, t["reshape_layer.weight"], t["reshape_layer.bias"])
    h = F.linear(torch.relu(F.linear(q, t["support_encoder.proj1.weight"], t["support_encoder.proj1.bias"])),
                 t["support_encoder.proj2.weight"], t["support_encoder.proj2.bias"])
    D = q.shape[-1]
    return F.layer_norm(h + q, (D,), t["support_encoder.layer_norm.weight"], t["support_encoder.layer_norm.bias"], 1e-5).numpy()


def cosine_mean_scores(cand_vecs, rel_vecs):
    """sklearn.metrics.pairwise.cosine_similarity(cand, rel).mean(axis=1) restated in numpy float32 (normalise rows, dot, mean)"""
    def norm(x):
        n = np.sqrt((x.astype(np.float32) ** 2).sum(1, dtype=np.float32))
        n[n == 0] = 1.0
        return (x / n[:, None]).astype(np.float32)
    return (norm(cand_vecs) @ norm(rel_vecs).T).mean(axis=1, dtype=np.float32)


def rank_interval(scores):
    """(optimistic, pessimistic) rank of candidate 0 under a descending sort: zsl_module.py:705-706's argsort leaves exact ties unpinned"""
    s0 = scores[0]
    gt = int((scores[1:] > s0).sum()); eq = int((scores[1:] == s0).sum())
    return gt + 1, gt + eq + 1


def entity_halves(w, ent_symbol, conn, deg):
    """The separable form: per entity A_e (its contribution as the pair's head) and B_e (as the candidate tail), so that
    reshape_layer([N_h | tanh fc1(h) | tanh fc2(c) | N_c]) = A_h + B_c.  float64 accumulation, returned as float32."""
    W = {k: v.astype(np.float64) for k, v in w.items()}
    emb = W["symbol_emb.weight"]
    D = emb.shape[1]; H = D // 2
    nb = emb[conn].sum(1) @ W["gcn_w.weight"].T + conn.shape[1] * W["gcn_w.bias"]
    with np.errstate(divide="ignore", invalid="ignore"):
        N = np.tanh(nb / np.asarray(deg, np.float64)[:, None])
    T1 = np.tanh(emb[ent_symbol] @ W["fc1.weight"].T + W["fc1.bias"])
    T2 = np.tanh(emb[ent_symbol] @ W["fc2.weight"].T + W["fc2.bias"])
    Wr = W["reshape_layer.weight"]
    A = N @ Wr[:, :H].T + T1 @ Wr[:, H:2 * H].T
    B = T2 @ Wr[:, 2 * H:3 * H].T + N @ Wr[:, 3 * H:].T + W["reshape_layer.bias"]
    return A.astype(np.float32), B.astype(np.float32)


def separable_scores(w, A, B, head, cands, rel_vecs):
    """scores of one candidate list from the entity halves (float64 inside): what the CUDA path computes"""
    W = {k: v.astype(np.float64) for k, v in w.items()}
    x = A[head].astype(np.float64)[None, :] + B[cands].astype(np.float64)
    h = np.maximum(x @ W["support_encoder.proj1.weight"].T + W["support_encoder.proj1.bias"], 0)
    y = h @ W["support_encoder.proj2.weight"].T + W["support_encoder.proj2.bias"] + x
    mu = y.mean(1, keepdims=True); var = y.var(1, keepdims=True)
    g = (y - mu) / np.sqrt(var + 1e-5) * W["support_encoder.layer_norm.weight"] + W["support_encoder.layer_norm.bias"]
    r = rel_vecs.astype(np.float64)
    cos = (g @ r.T) / (np.linalg.norm(g, axis=1)[:, None] * np.linalg.norm(r, axis=1)[None, :])
    return cos.mean(1)


def tf32_split(x):
    """hi + lo as the tensor-core kernel splits an FP32 operand: hi = round-to-nearest (ties away) to 11 significand bits, lo =
    the exact FP32 remainder cut to its leading 11 bits (what kind::tf32 reads of it)"""
    x = np.ascontiguousarray(x, np.float32)
    hi = ((x.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    lo = ((x - hi).astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    return hi, lo


def tensor_core_scores(w, A, B, head, cands, rel_vecs, split=True):
    """csrc/zsl_rank.cu:zsl_tc_kernel restated: hidden layer split per entity (A1 = W1 A + b1, B1 = W1 B, FP32), the 400 -> 200
    contraction as hi*hi + hi*lo + lo*hi of TF32-split operands (products and sums in float64 here: the tensor core adds FP32
    rounding on top), the LayerNorm mean from the column-sum row of W2, and the one-pass centred sums
        var = sum d^2 / D,  z . rsum = rstd sum (d g) rsum + be . rsum,  |z|^2 = rstd^2 sum (d g)^2 + 2 rstd sum (d g) be + |be|^2
    with rsum = sum_k r_k / |r_k| (the cosine mean is linear in the normalised relation vectors)."""
    f32 = np.float32
    W1, b1 = w["support_encoder.proj1.weight"].astype(f32), w["support_encoder.proj1.bias"].astype(f32)
    W2, b2 = w["support_encoder.proj2.weight"].astype(f32), w["support_encoder.proj2.bias"].astype(f32)
    g, be = w["support_encoder.layer_norm.weight"].astype(f32), w["support_encoder.layer_norm.bias"].astype(f32)
    D = W2.shape[0]
    a1 = (A[head].astype(np.float64) @ W1.T.astype(np.float64) + b1).astype(f32)
    b1c = (B[cands].astype(np.float64) @ W1.T.astype(np.float64)).astype(f32)
    hid = np.maximum(a1[None, :] + b1c, f32(0))                                        # FP32 add, as the gather warps do
    wsum = W2.sum(0, dtype=np.float32)                                                  # the extra output row: column sums
    W2x = np.concatenate([W2, wsum[None, :]], 0)
    if split:
        hh, hl = tf32_split(hid)
        wh = ((W2x.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(f32)
        r = (W2x - wh).astype(f32)
        wl = ((r.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(f32)
        acc = (hl.astype(np.float64) @ wh.T.astype(np.float64) + hh.astype(np.float64) @ wl.T.astype(np.float64)
               + hh.astype(np.float64) @ wh.T.astype(np.float64))
    else:
        acc = hid.astype(np.float64) @ W2x.T.astype(np.float64)
    acc = acc.astype(f32)
    x = A[head][None, :].astype(f32) + B[cands].astype(f32)
    mu = ((acc[:, D] + b2.sum(dtype=f32)) + (A[head].sum(dtype=f32) + B[cands].sum(1, dtype=f32))) / f32(D)
    u = (acc[:, :D] + b2[None, :]) + x
    d = (u - mu[:, None]).astype(np.float64)
    e = d * g
    rn = np.linalg.norm(rel_vecs.astype(np.float64), axis=1)
    rsum = (rel_vecs.astype(np.float64) / np.where(rn > 0, rn, 1.0)[:, None] * (rn > 0)[:, None]).sum(0)
    rstd = 1.0 / np.sqrt((d * d).sum(1) / D + 1e-5)
    dot = rstd * (e @ rsum) + float(be.astype(np.float64) @ rsum)
    nn = rstd * rstd * (e * e).sum(1) + 2.0 * rstd * (e @ be.astype(np.float64)) + float((be.astype(np.float64) ** 2).sum())
    return np.where(nn > 0, dot / np.sqrt(np.where(nn > 0, nn, 1.0)), 0.0) / rel_vecs.shape[0]
