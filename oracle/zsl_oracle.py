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
import numpy as np
import torch
import torch.nn.functional as F

D_DEFAULT = 200


def seeded_extractor_weights(seed, n_symbols, D=D_DEFAULT):
    """Extractor state (names as in the reference's state_dict) drawn with numpy PCG64: same bits on any box."""
    rng = np.random.default_rng(seed)
    def lin(o, i):
        a = np.sqrt(6.0 / (o + i))
        return rng.uniform(-a, a, (o, i)).astype(np.float32), rng.uniform(-0.1, 0.1, o).astype(np.float32)
    w = {}
    emb = (rng.standard_normal((n_symbols + 1, D)) / np.sqrt(D)).astype(np.float32)
    emb[n_symbols] = 0.0                                        # padding_idx row
    w["symbol_emb.weight"] = emb
    for name, (o, i) in {"gcn_w": (D // 2, D), "fc1": (D // 2, D), "fc2": (D // 2, D), "reshape_layer": (D, 2 * D),
                         "support_encoder.proj1": (2 * D, D), "support_encoder.proj2": (D, 2 * D)}.items():
        w[name + ".weight"], w[name + ".bias"] = lin(o, i)
    w["gcn_b"] = np.zeros(D, np.float32)                        # declared by the reference, never used in forward
    w["support_encoder.layer_norm.weight"] = (1.0 + 0.1 * rng.standard_normal(D)).astype(np.float32)
    w["support_encoder.layer_norm.bias"] = (0.1 * rng.standard_normal(D)).astype(np.float32)
    return w


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


def synthetic_zsl_setup(seed=192, n_ent=300, n_rel=3, D=D_DEFAULT, max_nb=50, T=14):
    """The seeded graph of tests/golden/golden_zsl.npz: symbols = entities 0..n_ent-1, then the relations, then the pad id;
    connections [n_ent, max_nb, 2] (relation symbol, neighbour symbol) padded with the pad id, degrees, T candidate lists."""
    rng = np.random.default_rng(seed)
    n_symbols = n_ent + n_rel
    deg = rng.integers(1, max_nb + 1, n_ent)
    deg[:5] = [1, 2, max_nb, max_nb, 3]
    conn = np.full((n_ent, max_nb, 2), n_symbols, np.int64)
    for e in range(n_ent):
        conn[e, :deg[e], 0] = n_ent + rng.integers(0, n_rel, deg[e])
        conn[e, :deg[e], 1] = rng.integers(0, n_ent, deg[e])
    sizes = [1, 2, 17, 33, 64, 65, 128, 129, 150, 200, 257, 40, 90, 7][:T]
    heads = rng.integers(0, n_ent, T)
    rels = rng.integers(0, n_rel, T)
    cands = [rng.choice(n_ent, s, replace=False).astype(np.int64) for s in sizes]      # candidate 0 = the true tail
    cands[3][5] = cands[3][0]                                                           # an exact tie with the true candidate
    rel_vecs = rng.standard_normal((n_rel, 20, D)).astype(np.float32)
    return n_symbols, conn, deg.astype(np.float32), heads, rels, cands, rel_vecs
