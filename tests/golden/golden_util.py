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


# ---- paper half (tests/golden/make_golden_paper.py, tests/test_paper_golden.py): seeded inputs of golden_paper.npz
PAPER_CALC_CASES = (
    # name, score_model, mode, score_norm_flag, B (relation rows), k (candidate blocks; 1 for 'normal')
    ("transe_normal", "transe", "normal", False, 96, 1),
    ("transe_normal_norm", "transe", "normal", True, 96, 1),
    ("transe_head_batch", "transe", "head_batch", False, 8, 40),
    ("transe_tail_batch", "transe", "tail_batch", False, 8, 40),
    ("distmult_normal", "distmult", "normal", False, 96, 1),
    ("distmult_head_batch", "distmult", "head_batch", False, 8, 40),
    ("distmult_tail_batch", "distmult", "tail_batch", False, 8, 40),
)


def paper_calc_inputs(seed, name, D=200):
    """(h, t, r) float32 row tensors of one _calc case, shaped as module/NegativeSampling.py:142-168 expects them:
    normal: all [B, D]; head_batch: h [k*B, D], t and r [B, D]; tail_batch: t [k*B, D], h and r [B, D]."""
    case = next(c for c in PAPER_CALC_CASES if c[0] == name)
    _, _, mode, _, B, k = case
    rng = np.random.default_rng([seed, PAPER_CALC_CASES.index(case)])
    draw = lambda n: (rng.standard_normal((n, D)) * 0.3).astype(np.float32)
    h = draw(k * B if mode == "head_batch" else B)
    t = draw(k * B if mode == "tail_batch" else B)
    r = draw(B)
    return h, t, r


def paper_subgraph_cases(seed=SEED, n_cases=4):
    """Seeded sampled subgraphs for neg_sample_fn (module/NegativeSampling.py:114-140): a global train graph (E = 120,
    R = 4, dense enough that the filter rejects often), then per case a node subset in LOCAL ids and the train edges inside it.
    Returns (whole_triples (h, r, t) lists of GLOBAL ids, [case dicts: l2g, edge_h, edge_t, edge_r, neg_ent, py_seed])."""
    rng = np.random.default_rng([seed, 77])
    E, R, n_train = 120, 4, 2600
    th, tt, tr = rng.integers(0, E, n_train), rng.integers(0, E, n_train), rng.integers(0, R, n_train)
    cases = []
    for c in range(n_cases):
        n_local = (30, 48, 64, 90)[c % 4]
        l2g = rng.choice(E, n_local, replace=False).astype(np.int64)
        g2l = {int(g): i for i, g in enumerate(l2g)}
        inside = np.array([i for i in range(n_train) if int(th[i]) in g2l and int(tt[i]) in g2l], np.int64)
        pick = rng.choice(inside, min((25, 60, 90, 120)[c % 4], len(inside)), replace=False)
        cases.append(dict(l2g=l2g, edge_h=np.array([g2l[int(th[i])] for i in pick], np.int64),
                          edge_t=np.array([g2l[int(tt[i])] for i in pick], np.int64), edge_r=tr[pick].astype(np.int64),
                          neg_ent=(1, 4, 8, 16)[c % 4], py_seed=1000 + c))
    return (th.tolist(), tr.tolist(), tt.tolist()), E, R, cases


def paper_eval_setup(seed=SEED, E=400, R=6, D=200, per_rel=15):
    """Seeded inputs of main.evaluate (main.py:217-272): entity / relation symbols and ids, embeddings, and a
    {mode}_candidates.json-shaped dict {relation: {"head\\trel\\ttail": [true tail, candidates...]}} with ragged candidate
    lists (1 ... 300) and planted EXACT ties (clones of the true tail's embedding row among the candidates)."""
    rng = np.random.default_rng([seed, 99])
    ents = [f"/m/e{i:03d}" for i in range(E)]
    rels = [f"/r/rel{i}" for i in range(R)]
    e2id = {e: i for i, e in enumerate(ents)}
    r2id = {r: i for i, r in enumerate(rels)}
    ent = (rng.standard_normal((E, D)) * 0.25).astype(np.float32)
    rel = (rng.standard_normal((R, D)) * 0.25).astype(np.float32)
    # entities 380..399 are clones in groups of five: rows identical -> exactly equal scores for any query
    for g in range(4):
        ent[380 + 5 * g: 385 + 5 * g] = ent[380 + 5 * g]
    sizes = [1, 2, 3, 17, 64, 65, 128, 129, 200, 257, 300, 33, 90, 150, 7]
    cand = {}
    for ri, rname in enumerate(rels):
        items = {}
        for k in range(per_rel):
            head = int(rng.integers(0, 380))
            n = sizes[(k + ri) % len(sizes)]
            if k % 5 == 0 and n >= 5:      # planted ties: the true tail is a clone, 1..4 of its siblings are candidates
                g = int(rng.integers(0, 4))
                true = 380 + 5 * g
                sib = [true + 1 + j for j in range(1 + (k // 5 + ri) % 4)]
                rest = rng.choice(380, n - 1 - len(sib), replace=False).tolist()
                lst = [true] + sib + rest
            else:
                lst = rng.choice(380, n, replace=False).tolist()
            if head in lst[1:]:
                lst = [x for x in lst if x != head] if lst[0] != head else lst
            key = "\t".join((ents[head], rname, ents[lst[0]]))
            if key in items:
                continue
            items[key] = [ents[i] for i in lst]
        cand[rname] = items
    return ents, rels, e2id, r2id, ent, rel, cand
