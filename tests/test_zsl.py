"""ZSL candidate scorer (ZSLmodule.eval, module/zsl_module.py:635-745 -> mre_zsl_entity_features + mre_zsl_rank).

Fixture tests/golden/golden_zsl.npz: scores and ranks produced by the REFERENCE's own Extractor class + sklearn
cosine_similarity on a seeded synthetic graph (tests/golden/make_golden_zsl.py asserts oracle/zsl_oracle.py restates the
Extractor bit for bit).  CPU: the oracle's separable form (what the kernels compute) against the reference scores.
GPU: entity halves and scores against the oracle (FP32 rounding: 2e-5 absolute on cosine means in [-1, 1]), ranks inside the
interval the reference's scores leave open at that tolerance and equal to the reference's rank wherever it is empty,
(hits10, hits5, mrr) within 1e-4 when no rank is in doubt.
"""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import zsl_oracle as zo

TOL = 1e-6


@pytest.fixture(scope="module")
def setup():
    g = gu.load("golden_zsl.npz")
    n_symbols, conn, deg, heads, rels, cands, rel_vecs = zo.synthetic_zsl_setup(int(g["seed"]), int(g["n_ent"]), int(g["n_rel"]),
                                                                                int(g["D"]), int(g["max_nb"]), len(g["ranks"]))
    w = zo.seeded_extractor_weights(int(g["seed"]), n_symbols, int(g["D"]))
    return g, w, conn, deg, heads, rels, cands, rel_vecs


def test_separable_form_matches_reference_scores(setup):
    g, w, conn, deg, heads, rels, cands, rel_vecs = setup
    n_ent = int(g["n_ent"])
    A, B = zo.entity_halves(w, np.arange(n_ent), conn[:, :, 1], deg)
    for t in range(len(cands)):
        ref = g["scores"][g["ptr"][t]:g["ptr"][t + 1]]
        mine = zo.separable_scores(w, A, B, int(heads[t]), cands[t], rel_vecs[rels[t]])
        assert np.abs(mine - ref).max() < TOL
        lo, hi = zo.rank_interval(ref)
        assert lo <= g["ranks"][t] <= hi
    # the planted exact tie with the true candidate leaves the reference's rank genuinely open
    lo, hi = zo.rank_interval(g["scores"][g["ptr"][3]:g["ptr"][4]])
    assert hi == lo + 1


def test_tensor_core_restatement_matches_reference_scores(setup):
    """oracle/zsl_oracle.py:tensor_core_scores restates what zsl_tc_kernel computes -- per-entity hidden halves, the TF32
    hi/lo split with the lo*lo product dropped, the LayerNorm mean from a column-sum row of W2, one-pass centred sums, the
    cosine mean folded into one normalised relation sum -- and must land on the reference's own scores"""
    g, w, conn, deg, heads, rels, cands, rel_vecs = setup
    n_ent = int(g["n_ent"])
    A, B = zo.entity_halves(w, np.arange(n_ent), conn[:, :, 1], deg)
    worst = 0.0
    for t in range(len(cands)):
        ref = g["scores"][g["ptr"][t]:g["ptr"][t + 1]]
        mine = zo.tensor_core_scores(w, A, B, int(heads[t]), cands[t], rel_vecs[rels[t]])
        worst = max(worst, float(np.abs(mine - ref).max()))
        # the split costs nothing measurable: the same algebra with unsplit FP32 operands is no closer
        full = zo.tensor_core_scores(w, A, B, int(heads[t]), cands[t], rel_vecs[rels[t]], split=False)
        assert np.abs(mine - full).max() < 2e-7
    assert worst < 2e-7, worst
    # the split itself: hi + lo reproduces an FP32 value to 2^-21 relative, hi has 11 significand bits
    x = np.random.default_rng(0).standard_normal(4096).astype(np.float32)
    x = np.abs(x)
    hi, lo = zo.tf32_split(x)
    assert (hi.view(np.uint32) & 0x1FFF).max() == 0 and (lo.view(np.uint32) & 0x1FFF).max() == 0
    assert (np.abs((hi.astype(np.float64) + lo) - x) <= np.abs(x) * 2.0 ** -21).all()


def rank_bounds(ref_scores, tol):
    """ranks the reference scores allow when every score may move by tol"""
    s0 = ref_scores[0]
    lo = int((ref_scores[1:] > s0 + 2 * tol).sum()) + 1
    hi = int((ref_scores[1:] >= s0 - 2 * tol).sum()) + 1
    return lo, hi


@pytest.mark.gpu
def test_zsl_kernels_vs_reference(mre, setup):
    g, w, conn, deg, heads, rels, cands, rel_vecs = setup
    n_ent = int(g["n_ent"])
    ev = mre.paper.ZSLEvaluator(w, conn, deg, np.arange(n_ent), device=0)
    A, B = zo.entity_halves(w, np.arange(n_ent), conn[:, :, 1], deg)
    assert np.abs(ev.A.cpu().numpy() - A).max() < 1e-5 and np.abs(ev.B.cpu().numpy() - B).max() < 1e-5
    counts, scores = ev.rank(heads, rels, cands, rel_vecs, want_scores=True)
    c, s = counts.cpu().numpy(), scores.cpu().numpy()
    err = float(np.abs(s - g["scores"]).max())
    print(f"max |score - reference score| = {err:.3e}")
    assert err < TOL                                                 # against the reference's own scores
    for t in (2, 9):                                                 # and against the oracle's restatement of the tensor-core algebra
        want = zo.tensor_core_scores(w, A, B, int(heads[t]), cands[t], rel_vecs[rels[t]])
        assert np.abs(s[g["ptr"][t]:g["ptr"][t + 1]] - want).max() < TOL
    exact = 0
    for t in range(len(cands)):
        ref = g["scores"][g["ptr"][t]:g["ptr"][t + 1]]
        lo, hi = rank_bounds(ref, err + 1e-7)                        # what the reference's scores allow at the measured deviation
        assert lo <= c[0][t] + 1 and c[0][t] + c[1][t] + 1 <= hi, t
        if lo == hi:
            exact += 1
            assert c[0][t] + 1 == g["ranks"][t] and c[1][t] == 0
        # the kernel's own counts are consistent with its own scores
        mine = s[g["ptr"][t]:g["ptr"][t + 1]]
        assert (c[0][t], c[1][t]) == (int((mine[1:] > mine[0]).sum()), int((mine[1:] == mine[0]).sum()))
    assert exact >= len(cands) - 3
    assert c[1][3] >= 1                                              # the planted duplicate candidate ties exactly
    # metrics through mre_metrics (pessimistic = a stable descending sort's answer)
    rk = mre.engine.Ranker(ev.ctx)
    m = rk.metrics(counts, 1, "pessimistic")
    sums, rr = m["sums"].cpu().numpy(), m["rr"].cpu().numpy()
    ranks = (c[0] + c[1] + 1).astype(np.float64)
    assert sums[1][0] == len(cands) and np.isclose(rr[1], (1.0 / ranks).sum(), rtol=1e-12)
    assert sums[1][5] == (ranks <= 10).sum() and sums[1][4] == (ranks <= 5).sum()


@pytest.mark.gpu
def test_zsl_eval_dropin(mre, setup, capsys):
    """ZSLmodule.eval's contract: {relation: {"head\\trel\\ttail": [true, cands...]}} in, (hits10, hits5, mrr) out"""
    g, w, conn, deg, heads, rels, cands, rel_vecs = setup
    n_ent = int(g["n_ent"])
    ent2id = {f"e{i}": i for i in range(n_ent)}
    order = np.argsort(rels, kind="stable")
    test_candidates, relation_vecs = {}, {}
    for t in order.tolist():
        rel = f"r{int(rels[t])}"
        key = "\t".join((f"e{int(heads[t])}", rel, f"e{int(cands[t][0])}")) + f"#{t}"      # keys must be unique per triple
        test_candidates.setdefault(rel, {})[key] = [f"e{int(x)}" for x in cands[t]]
        relation_vecs[rel] = rel_vecs[rels[t]]
    ev = mre.paper.ZSLEvaluator(w, conn, deg, np.arange(n_ent), device=0)
    h10, h5, mrr = ev.eval(test_candidates, relation_vecs, ent2id, mode="test", ties="optimistic")
    out = capsys.readouterr().out
    assert "HITS10" in out and "MAP" in out
    ref = g["metrics"]
    # one triple carries a planted exact tie: allow its rank to differ by one
    assert abs(h10 - ref[0]) <= 1 / len(cands) + 1e-9 and abs(h5 - ref[1]) <= 1 / len(cands) + 1e-9 and abs(mrr - ref[2]) < 5e-3


@pytest.mark.gpu
def test_zsl_ragged_and_empty(mre, setup):
    g, w, conn, deg, heads, rels, cands, rel_vecs = setup
    n_ent = int(g["n_ent"])
    ev = mre.paper.ZSLEvaluator(w, conn, deg, np.arange(n_ent), device=0)
    counts, scores = ev.rank(np.zeros(0, np.int64), np.zeros(0, np.int64), [], rel_vecs, want_scores=True)
    assert counts.shape == (4, 0)
    # a list holding only the true candidate ranks first; a longer sweep crosses the 64-row tile edge mid-list
    lists = [cands[0], np.arange(n_ent, dtype=np.int64), cands[1]]
    counts, scores = ev.rank(heads[:3], rels[:3], lists, rel_vecs, want_scores=True)
    c = counts.cpu().numpy()
    assert c[0][0] == 0 and c[1][0] == 0
    A, B = zo.entity_halves(w, np.arange(n_ent), conn[:, :, 1], deg)
    want = zo.separable_scores(w, A, B, int(heads[1]), lists[1], rel_vecs[rels[1]])
    got = scores.cpu().numpy()[1:1 + n_ent]
    assert np.abs(got - want).max() < TOL


@pytest.mark.gpu
def test_zsl_tensor_core_path_vs_fp32_path_fullsize(mre, monkeypatch):
    """FB15K-237-ZS-sized sweep (14 208 entities, 2 000 triples x 1 000 candidates = 2 M pairs, tiles straddling triples and
    a ragged tail): the 3xTF32 tensor-core kernel against the FP32 CUDA-core kernels of the same library on every pair
    (scores within 1e-6, the north-star band is 1e-5), each path's counts consistent with its own scores, and ranks equal
    wherever the FP32 scores leave no candidate within 2e-6 of the true one."""
    E, R, D, NB, T, C = 14208, 29, 200, 50, 2000, 1000
    rng = np.random.default_rng(5)
    w = gu.seeded_extractor_weights(3, E + R, D)
    conn = rng.integers(0, E, (E, NB)).astype(np.int64)
    deg = rng.integers(1, NB + 1, E).astype(np.float32)
    ev = mre.paper.ZSLEvaluator(w, conn, deg, np.arange(E), device=0)
    heads, rels = rng.integers(0, E, T), rng.integers(0, R, T)
    cands = [rng.choice(E, C - (t % 7), replace=False) for t in range(T)]      # ragged lists: tiles straddle triples
    rel_vecs = rng.standard_normal((R, 20, D)).astype(np.float32)
    ev.ctx.option("zsl_fp32", 1)
    c32, s32 = ev.rank(heads, rels, cands, rel_vecs, want_scores=True)
    ev.ctx.option("zsl_fp32", 0)
    ctc, stc = ev.rank(heads, rels, cands, rel_vecs, want_scores=True)
    s32, stc, c32, ctc = s32.cpu().numpy(), stc.cpu().numpy(), c32.cpu().numpy(), ctc.cpu().numpy()
    err = float(np.abs(s32 - stc).max())
    print(f"tensor-core vs FP32 path: max |score diff| = {err:.3e} over {len(stc)} pairs")
    assert err < TOL
    ptr = np.concatenate([[0], np.cumsum([len(c) for c in cands])])
    # against the reference: oracle/zsl_oracle.py restates the reference's Extractor + cosine mean BIT FOR BIT (pinned by
    # tests/golden/make_golden_zsl.py); a spread of triples of THIS full-size sweep, every candidate of each
    from oracle import zsl_oracle as zo
    worst = 0.0
    for t in np.linspace(0, T - 1, 12).astype(np.int64).tolist():
        c = cands[t]
        left = np.full(len(c), heads[t])
        vecs = zo.extractor_query_vectors(w, np.stack([left, c], 1), conn[left], deg[left], conn[c], deg[c])
        ref = zo.cosine_mean_scores(vecs, rel_vecs[rels[t]])
        worst = max(worst, float(np.abs(stc[ptr[t]:ptr[t + 1]] - ref).max()))
        lo, hi = zo.rank_interval(ref)                      # the reference's argsort rank lies in [lo, hi] (exact ties unpinned)
        if np.abs(ref[1:] - ref[0]).min() > 2e-6:
            assert lo - 1 <= ctc[0][t] <= hi - 1
    print(f"tensor-core path vs the pinned reference restatement: max |score diff| = {worst:.3e}")
    assert worst < TOL
    same = 0
    for t in range(T):
        a, b = s32[ptr[t]:ptr[t + 1]], stc[ptr[t]:ptr[t + 1]]
        assert (ctc[0][t], ctc[1][t]) == (int((b[1:] > b[0]).sum()), int((b[1:] == b[0]).sum()))
        assert (c32[0][t], c32[1][t]) == (int((a[1:] > a[0]).sum()), int((a[1:] == a[0]).sum()))
        if np.abs(a[1:] - a[0]).min() > 2e-6:
            same += 1
            assert ctc[0][t] == c32[0][t] and ctc[1][t] == 0
    assert same > 0.5 * T                                            # (1 000 scores in a narrow range: a fifth of the triples have a near-tie)


@pytest.mark.gpu
@pytest.mark.parametrize("D", [8, 104, 216])
def test_zsl_tensor_core_path_other_dims(mre, monkeypatch, D):
    """model dimensions other than the reference's 200 (odd chunk counts, one k-block, the widest tile that fits): the
    tensor-core kernel against the FP32 CUDA-core kernels, ragged lists, a tile tail and a single-candidate list"""
    E, R, NB, T = 700, 5, 12, 9
    rng = np.random.default_rng(D)
    w = gu.seeded_extractor_weights(11, E + R, D)
    conn = rng.integers(0, E, (E, NB)).astype(np.int64)
    deg = rng.integers(1, NB + 1, E).astype(np.float32)
    ev = mre.paper.ZSLEvaluator(w, conn, deg, np.arange(E), device=0)
    heads, rels = rng.integers(0, E, T), rng.integers(0, R, T)
    sizes = [1, 300, 129, 127, 256, 257, 3, 511, 64]
    cands = [rng.choice(E, s, replace=False) for s in sizes]
    rel_vecs = rng.standard_normal((R, 20, D)).astype(np.float32)
    ev.ctx.option("zsl_fp32", 1)
    c32, s32 = ev.rank(heads, rels, cands, rel_vecs, want_scores=True)
    ev.ctx.option("zsl_fp32", 0)
    ctc, stc = ev.rank(heads, rels, cands, rel_vecs, want_scores=True)
    s32, stc, ctc = s32.cpu().numpy(), stc.cpu().numpy(), ctc.cpu().numpy()
    from oracle import zsl_oracle as zo                      # and against the pinned restatement of the reference Extractor
    for t in range(T):
        c = cands[t]
        left = np.full(len(c), heads[t])
        vecs = zo.extractor_query_vectors(w, np.stack([left, c], 1), conn[left], deg[left], conn[c], deg[c])
        off = int(np.sum(sizes[:t]))
        assert np.abs(stc[off:off + len(c)] - zo.cosine_mean_scores(vecs, rel_vecs[rels[t]])).max() < TOL
    assert np.isfinite(stc).all() and np.abs(s32 - stc).max() < TOL
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    for t in range(T):
        b = stc[ptr[t]:ptr[t + 1]]
        assert (ctc[0][t], ctc[1][t]) == (int((b[1:] > b[0]).sum()), int((b[1:] == b[0]).sum()))
