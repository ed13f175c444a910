"""Shared by tests/golden/make_golden.py (build container, reads /root/reference) and the tests
(any box, reads only the committed .npz fixtures)."""
import os

import numpy as np

GOLDEN_DIR = os.path.dirname(os.path.abspath(__file__))
SEED = 192  # the reference's default seed (args.py:8)
TIE_BAND = 1e-5  # north_star: ranks are exact for score gaps > 1e-5 relative


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name))


def xavier_tables(seed, shapes):
    """Seeded xavier-uniform float32 tables (the init OpenKE uses, TransE.py:21-22), drawn with numpy's
    PCG64 so the same bits come back on any box."""
    rng = np.random.default_rng(seed)
    out = []
    for rows, dim in shapes:
        a = np.sqrt(6.0 / (rows + dim))
        out.append(rng.uniform(-a, a, (rows, dim)).astype(np.float32))
    return out


def structured_tables(seed, shapes, rank=8):
    """Seeded low-rank-plus-noise float32 tables: scores spread over a wide range (as a trained model's
    do), so almost no candidate falls inside the 1e-5 tie band and exact rank equality is demanded."""
    rng = np.random.default_rng(seed + 1)
    out = []
    for rows, dim in shapes:
        z = rng.standard_normal((rows, rank))
        w = rng.standard_normal((rank, dim))
        x = np.zeros((rows, dim))
        for k in range(rank):  # explicit order: no BLAS, same bits on any CPU
            x += z[:, k:k + 1] * w[k:k + 1, :]
        x = x / np.sqrt(rank) + 0.1 * rng.standard_normal((rows, dim))
        out.append((x / np.sqrt(dim)).astype(np.float32))
    return out


WEIGHT_SETS = {"xavier": xavier_tables, "structured": structured_tables}


def split_cols(z, prefix):
    return tuple(z[f"{prefix}_{c}"].astype(np.int64) for c in "htr")


def group_lists(key_a, key_b, val):
    """dict (a, b) -> sorted unique array of val."""
    out = {}
    for a, b, v in zip(key_a.tolist(), key_b.tolist(), val.tolist()):
        out.setdefault((a, b), set()).add(v)
    return {k: np.fromiter(sorted(v), np.int64, len(v)) for k, v in out.items()}


def band_counts(scores, true_idx, known, band):
    """Filtered strict-rank interval implied by a float32 score vector under a tie band:
    lo = #{j unfiltered : s_j < s_true - band},  hi = #{j unfiltered : s_j <= s_true + band}."""
    s_true = scores[true_idx]
    mask = np.ones(len(scores), bool)
    mask[known] = False
    mask[true_idx] = False
    s = scores[mask]
    return int((s < s_true - band).sum()), int((s <= s_true + band).sum())


# ---- ZSL scorer inputs (tests/test_zsl.py, tests/golden/make_golden_zsl.py, bench_zsl.py)
def seeded_extractor_weights(seed, n_symbols, D=200):
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


def synthetic_zsl_setup(seed=192, n_ent=300, n_rel=3, D=200, max_nb=50, T=14):
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
