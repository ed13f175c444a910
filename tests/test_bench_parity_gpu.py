"""Parity of the EXACT workloads bench.py runs (bench.load_workload: same ids, same seeded tables, same call through
bench.Runner.prepare -> mre_rank) against

  * oracle/kge_oracle.c (sequential FP32): raw and filtered counts BIT-EXACT on >= 500 sampled queries per workload
    (TransE; DistMult; ComplEx in contraction form -- the association of the tcgen05 path's exact re-score);
  * tests/golden/golden_bench.npz -- the REAL reference on the same workloads (reference OpenKE modules -> unmodified
    Base.so testTail; reference NegativeSampling.evaluate + main.evaluate for the candidate path): filtered counts inside
    the north_star's 1e-5 relative tie band, equal wherever the band is empty, metric tuple within 1e-4;
  * the synthetic 2 M x 256 table (BASELINE configs[4]) against the C oracle's scores on sampled queries (tests/test_fullsize_gpu.py).
"""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import kge_oracle as ko

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def runner(mre):
    import bench
    return bench, bench.Runner(0, 1, 0)


def run_counts(runner, name, filt="index"):
    bench, R = runner
    w = bench.load_workload(name)
    w.filter = filt                 # "index": the known triples held by an mre_index (bench default); "csr": per-query lists
    p = R.prepare(w)
    p["step_dev"]()
    torch.cuda.synchronize()
    return w, p, p["counts_d"].cpu().numpy()


def oracle_scores(w, h, t, r):
    tabs = w.tables
    if w.scorer == "transe":
        return ko.transe_scores(tabs[0], tabs[1], w.p_norm, 1, h, t, r)
    if w.scorer == "distmult":
        return ko.distmult_scores(tabs[0], tabs[1], 1, h, t, r)
    return ko.complex_scores_contracted(tabs[0], tabs[1], tabs[2], tabs[3], 1, h, t, r)


@pytest.mark.parametrize("filt", ["index", "csr"])
@pytest.mark.parametrize("name", ["db15k_zs", "distmult", "complex"])
def test_all_entity_workloads_bit_exact_vs_sequential_oracle(runner, name, filt):
    w, p, c = run_counts(runner, name, filt)
    ptr, idx = w.filt_csr
    sel = np.unique(np.linspace(0, len(w.q_h) - 1, 520).astype(np.int64))
    assert len(sel) >= 500
    for i in sel.tolist():
        h, t, r = int(w.q_h[i]), int(w.q_t[i]), int(w.q_r[i])
        s = oracle_scores(w, h, t, r)
        known = idx[ptr[i]:ptr[i + 1]]
        raw = int((s < s[t]).sum())
        filt = raw - int((s[known[known != t]] < s[t]).sum())
        eq = int((s == s[t]).sum())                                # raw_eq counts the true entity itself (include/mre_b200.h)
        feq = eq - 1 - int((s[known[known != t]] == s[t]).sum())
        assert (c[0][i], c[1][i], c[2][i], c[3][i]) == (raw, eq, filt, feq), (name, i, c[:, i], raw, eq, filt, feq)


@pytest.mark.parametrize("name", ["db15k_zs", "distmult", "complex"])
def test_all_entity_workloads_vs_reference_golden(runner, name):
    """the reference's own modules + Base.so on the bench workload: inside the 1e-5 band, equal where it is empty"""
    bench, R = runner
    w, p, c = run_counts(runner, name)
    g = gu.load("golden_bench.npz")
    pos = {}
    for i, key in enumerate(zip(w.q_h.tolist(), w.q_r.tolist(), w.q_t.tolist())):
        pos.setdefault(key, i)
    i = np.array([pos[tuple(q)] for q in g[f"{name}_q"].tolist()])
    assert len(i) >= 500
    filt, raw = c[2][i], c[0][i]
    lo, hi, ref = g[f"{name}_lo"], g[f"{name}_hi"], g[f"{name}_filt"]
    outside = (filt < lo) | (filt > hi)
    assert not outside.any(), (name, int(outside.sum()), filt[outside][:5], lo[outside][:5], hi[outside][:5])
    exact = lo == hi
    assert np.array_equal(filt[exact], ref[exact])
    # Base.so's metric tuple for the same sample (tail side only; each side / testTotal, then (head + tail) / 2, Test.h:232-277)
    sums = bench.host_metric_sums(c[:, i], 1, "strict")
    T = float(g[f"{name}_test_total"])
    mine = np.array([sums[1][6] / 2.0 ** 32 / T / 2, sums[1][1] / T / 2, sums[1][5] / T / 2, sums[1][3] / T / 2, sums[1][2] / T / 2])
    ref_t = g[f"{name}_tuple"].astype(np.float64)
    assert np.allclose(mine[[0, 2, 3, 4]], ref_t[[0, 2, 3, 4]], atol=1e-4)
    assert np.isclose(mine[1], ref_t[1], rtol=1e-4)                      # MR: ranks moved only inside the band
    par = bench.parity_vs_golden(name, p, w)
    assert par["checked"] == len(i) and par["mismatches"] == 0


@pytest.mark.parametrize("filt", ["index", "csr"])
def test_candidate_workload_vs_reference_main_evaluate(runner, filt):
    """configs[0]: rel2candidates + ties//2 -- the reference class's evaluate + main.evaluate on the same candidate lists"""
    bench, R = runner
    w, p, c = run_counts(runner, "fb15k237_zs", filt)
    g = gu.load("golden_bench.npz")
    pos = {}
    for i, key in enumerate(zip(w.q_h.tolist(), w.q_r.tolist(), w.q_t.tolist())):
        pos.setdefault(key, i)
    i = np.array([pos[tuple(q)] for q in g["fb15k237_zs_q"].tolist()])
    assert len(i) >= 400
    mine = c[2][i].astype(np.int64) + c[3][i] // 2 + 1
    ref, band = g["fb15k237_zs_rank"], g["fb15k237_zs_band"]
    assert np.all(np.abs(mine - ref) <= band)
    assert np.array_equal(mine[band == 0], ref[band == 0]) and int((band == 0).sum()) >= 200
    fin = g["fb15k237_zs_final"]
    assert abs((1.0 / mine).mean() - fin[0]) <= 1e-4
    assert np.allclose([(mine <= k).mean() for k in (1, 3, 10)], fin[1:], atol=1e-4)
    # and bit for bit against the sequential oracle on the same candidate lists (utils/gen_mode_candidates.py:15-39)
    from oracle import paper_oracle as po
    z = gu.load("fb15k237_zs.npz")
    rel2cand = {int(rr): z["cand_ent"][k].astype(np.int64) for k, rr in enumerate(z["cand_rel"])}
    known = po.known_tails(w.q_h, w.q_r, w.q_t)
    sel = np.unique(np.linspace(0, len(w.q_h) - 1, 520).astype(np.int64))
    cands = po.build_candidates(w.q_h[sel], w.q_r[sel], w.q_t[sel], rel2cand, known)
    for k, q in enumerate(sel.tolist()):
        s = ko.transe_scores(w.tables[0], w.tables[1], 1, 1, int(w.q_h[q]), int(w.q_t[q]), int(w.q_r[q]))[cands[k]]
        assert po.rank_ties_half(s) == c[2][q] + c[3][q] // 2 + 1
    assert bench.parity_vs_golden("fb15k237_zs", p, w)["mismatches"] == 0


def test_openke_protocol_workload_counts(runner):
    """bench's fb15k237 workload (head + tail queries, normalised TransE, index filter) bit-exact vs the oracle on a sample"""
    bench, R = runner
    w, p, c = run_counts(runner, "fb15k237")
    import helpers
    ds = helpers.load_fb15k237()
    ent, rel = ko.l2_normalize_rows(w.tables[0]), ko.l2_normalize_rows(w.tables[1])
    q_h, q_t, q_r = p["q"]
    for i in np.linspace(0, len(q_h) - 1, 300).astype(np.int64).tolist():
        side = i % 2
        s = ko.transe_scores(ent, rel, 1, side, int(q_h[i]), int(q_t[i]), int(q_r[i]))
        raw, filt = ds.oracle.rank_from_scores(s, side, int(q_h[i]), int(q_t[i]), int(q_r[i]))
        assert (c[0][i], c[2][i]) == (raw, filt)


def test_sharded_blocks_add_up_to_the_whole_run(runner):
    """bench's strong-scaling cut (DistContext.shard_of blocks + sliced known-true CSR): the integer metric sums of 1, 2, 4, 8 ...
    blocks add up to exactly the single-process sums, fixed-point reciprocal-rank sum included -> bit-identical tuple at any N"""
    bench, R = runner
    w = bench.load_workload("db15k_zs")
    whole = R.prepare(w)["step_dev"]().cpu().numpy()
    for world in (2, 3, 8):
        tot = np.zeros_like(whole)
        for rank in range(world):
            R.rank, R.world = rank, world
            try:
                p = R.prepare(w)                       # this rank's block: sliced queries + sliced known-true CSR
                c = p["rank_dev"]()
                tot += R.rk.metrics(c, 1, "strict")["sums"].cpu().numpy()
            finally:
                R.rank, R.world = 0, 1
        assert np.array_equal(tot, whole), world
        assert R.eng.summarize(tot) == R.eng.summarize(whole)
