"""GPU parity of the fused score+rank path (TransE) against the oracle and the reference goldens.

Every call goes through the C ABI (mre_rank / mre_rank_host / mre_predict / mre_metrics).  Bars:
  * vs oracle/kge_oracle.c (same sequential-d float32 accumulation): raw and filtered counts BIT-EXACT;
  * vs the real reference (torch-CPU Model.predict -> Base.so testHead/testTail, tests/golden): counts inside the
    1e-5 relative tie band of SURVEY Appendix F and exactly equal wherever the band is empty; metric tuple 1e-4.
"""
import numpy as np
import pytest
import torch

import golden_util as gu
import helpers
from oracle import kge_oracle as ko, paper_oracle as po

pytestmark = pytest.mark.gpu

CONFIGS = {
    "transe_l1_norm": dict(p_norm=1, normalize=True),
    "transe_l2_norm": dict(p_norm=2, normalize=True),
    "transe_l1_raw": dict(p_norm=1, normalize=False),
}


@pytest.fixture(scope="module")
def env(mre, fb15k237):
    eng = mre.engine
    ix = eng.KGIndex.from_arrays(fb15k237.E, fb15k237.R, fb15k237.train, fb15k237.valid, fb15k237.test).to_device(0)
    rk = eng.Ranker(device=0)
    return eng, ix, rk


def both_sides(h, t, r):
    """queries in Tester order: (head query, tail query) per test triple"""
    q_h = np.repeat(h, 2); q_t = np.repeat(t, 2); q_r = np.repeat(r, 2)
    side = np.tile(np.array([0, 1], np.uint8), len(h))
    return q_h, q_t, q_r, side


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name", list(CONFIGS))
def test_transe_counts_bit_exact_vs_oracle(env, fb15k237, name):
    eng, ix, rk = env
    cfg = CONFIGS[name]
    E, R, D = fb15k237.E, fb15k237.R, 200
    ent, rel = gu.xavier_tables(gu.SEED, [(E, D), (R, D)])
    th, tt, tr = fb15k237.oracle.test_triples()
    sel = np.linspace(0, len(th) - 1, 150).astype(np.int64)
    q_h, q_t, q_r, side = both_sides(th[sel], tt[sel], tr[sel])
    ent_o, rel_o = (ko.l2_normalize_rows(ent), ko.l2_normalize_rows(rel)) if cfg["normalize"] else (ent, rel)
    raw_o, filt_o = helpers.oracle_counts(
        fb15k237, lambda s, h, t, r: ko.transe_scores(ent_o, rel_o, cfg["p_norm"], s, h, t, r), q_h, q_t, q_r, side)
    counts = rk.rank("transe", (dev(ent), dev(rel)), dev(q_h), dev(q_t), dev(q_r), dev(side), index=ix, **cfg).cpu().numpy()
    assert np.array_equal(counts[0], raw_o)
    assert np.array_equal(counts[2], filt_o)
    # the host-buffer entry point gives the same counts
    counts_h = rk.rank_host("transe", (dev(ent), dev(rel)), q_h, q_t, q_r, side, index=ix, **cfg)
    assert np.array_equal(counts_h, counts)


@pytest.mark.parametrize("wname", list(gu.WEIGHT_SETS))
@pytest.mark.parametrize("name", list(CONFIGS))
def test_transe_vs_reference_golden(env, fb15k237, name, wname):
    eng, ix, rk = env
    cfg = CONFIGS[name]
    g = gu.load("golden_fb15k237.npz")
    E, R, D = fb15k237.E, fb15k237.R, int(g["D"])
    ent, rel, _, _ = gu.WEIGHT_SETS[wname](gu.SEED, [(E, D), (R, D), (E, D), (R, D)])
    th, tt, tr = fb15k237.oracle.test_triples()
    qidx = g["qidx"]
    q_h, q_t, q_r, side = both_sides(th[qidx], tt[qidx], tr[qidx])
    tables = (dev(ent), dev(rel))
    counts = rk.rank("transe", tables, dev(q_h), dev(q_t), dev(q_r), dev(side), index=ix, **cfg)
    c = counts.cpu().numpy()
    key = f"{wname}_{name}"
    filt = c[2].reshape(-1, 2)
    lo, hi, ref = g[key + "_lo"], g[key + "_hi"], g[key + "_filt"]
    assert np.all(filt >= lo) and np.all(filt <= hi)
    exact = lo == hi
    assert np.array_equal(filt[exact], ref[exact])
    # metric tuple (mrr, mr, hit10, hit3, hit1), normalised by testTotal as the reference does
    m = rk.metrics(counts, dev(side), "strict")
    sums, rr = m["sums"].cpu().numpy(), m["rr"].cpu().numpy()
    T = float(fb15k237.oracle.test_total)
    mine = np.array([(rr[0] + rr[1]) / 2 / T, (sums[0][1] + sums[1][1]) / 2 / T, (sums[0][5] + sums[1][5]) / 2 / T,
                     (sums[0][3] + sums[1][3]) / 2 / T, (sums[0][2] + sums[1][2]) / 2 / T])
    ref_t = g[key + "_tuple"].astype(np.float64)
    assert np.allclose(mine[[0, 2, 3, 4]], ref_t[[0, 2, 3, 4]], atol=1e-4)
    assert np.isclose(mine[1], ref_t[1], rtol=2e-3)
    # Model.predict parity on the probe entities: 1e-5 relative to the score scale
    probe = g["probe"]
    for k in (0, len(qidx) // 2, len(qidx) - 1):
        for s in (0, 1):
            sc = rk.predict("transe", tables, dev(q_h), dev(q_t), dev(q_r), dev(side), query=2 * k + s, **cfg).cpu().numpy()
            refp = g[key + "_probe_scores"][k, s]
            assert np.allclose(sc[probe], refp, rtol=1e-5, atol=1e-6)


def test_predict_bit_exact_vs_oracle(env, fb15k237):
    eng, ix, rk = env
    E, R, D = fb15k237.E, fb15k237.R, 200
    ent, rel = gu.xavier_tables(7, [(E, D), (R, D)])
    q_h, q_t, q_r = np.array([5, 77]), np.array([900, 12000]), np.array([3, 200])
    side = np.array([0, 1], np.uint8)
    for p in (1, 2):
        for norm in (False, True):
            eo, ro = (ko.l2_normalize_rows(ent), ko.l2_normalize_rows(rel)) if norm else (ent, rel)
            for q in range(2):
                sc = rk.predict("transe", (dev(ent), dev(rel)), dev(q_h), dev(q_t), dev(q_r), dev(side), query=q, p_norm=p,
                                normalize=norm).cpu().numpy()
                want = ko.transe_scores(eo, ro, p, int(side[q]), int(q_h[q]), int(q_t[q]), int(q_r[q]))
                assert np.array_equal(sc, want)


@pytest.mark.parametrize("E,D,Q", [(1, 4, 1), (127, 8, 3), (129, 36, 130), (300, 6, 257), (1000, 200, 5), (513, 132, 64)])
def test_transe_ragged_shapes(mre, E, D, Q):
    """tile edges: E, Q not multiples of 128; D with a partial 32-float chunk; D not a multiple of 4 (padded)"""
    eng = mre.engine
    rng = np.random.default_rng(E * 1000 + D)
    R = 5
    ds = helpers.synthetic_graph(3, E, R, 4 * E, E // 2 + 1, Q)
    ix = eng.KGIndex.from_arrays(E, R, ds.train, ds.valid, ds.test).to_device(0)
    rk = eng.Ranker(device=0)
    ent = rng.standard_normal((E, D)).astype(np.float32)
    rel = rng.standard_normal((R, D)).astype(np.float32)
    th, tt, tr = ds.oracle.test_triples()
    side = (np.arange(Q) % 2).astype(np.uint8)
    for p in (1, 2):
        raw_o, filt_o = helpers.oracle_counts(ds, lambda s, h, t, r: ko.transe_scores(ent, rel, p, s, h, t, r), th, tt, tr, side)
        c = rk.rank("transe", (dev(ent), dev(rel)), dev(th), dev(tt), dev(tr), dev(side), index=ix, p_norm=p).cpu().numpy()
        assert np.array_equal(c[0], raw_o)
        assert np.array_equal(c[2], filt_o)


def test_empty_query_set(mre):
    eng = mre.engine
    rk = eng.Ranker(device=0)
    ent, rel = torch.randn(10, 8, device="cuda"), torch.randn(2, 8, device="cuda")
    z = torch.zeros(0, dtype=torch.int64, device="cuda")
    c = rk.rank("transe", (ent, rel), z, z, z, 1)
    assert c.shape == (4, 0)


def test_ties_and_rank_modes(mre):
    """duplicate entity rows give exact score ties: strict / ties_half / pessimistic conventions (main.py:245-250)"""
    eng = mre.engine
    rk = eng.Ranker(device=0)
    rng = np.random.default_rng(0)
    E, D, R = 64, 16, 2
    ent = rng.standard_normal((E, D)).astype(np.float32)
    ent[10:20] = ent[3]           # ten clones of entity 3
    rel = rng.standard_normal((R, D)).astype(np.float32)
    q_h, q_t, q_r = np.array([0]), np.array([3]), np.array([1])
    c = rk.rank("transe", (dev(ent), dev(rel)), dev(q_h), dev(q_t), dev(q_r), 1)
    s = ko.transe_scores(ent, rel, 1, 1, 0, 3, 1)
    lt = int((s < s[3]).sum()); eq = int((s == s[3]).sum()) - 1
    got = c.cpu().numpy()[:, 0]
    assert got[0] == lt and got[2] == lt and got[3] == eq and eq >= 10
    m = {k: rk.metrics(c, 1, k)["sums"].cpu().numpy()[1][1] for k in ("strict", "ties_half", "pessimistic")}
    assert m["strict"] == lt + 1 and m["ties_half"] == lt + eq // 2 + 1 and m["pessimistic"] == lt + eq + 1


def test_candidate_groups_paper_eval(mre):
    """FB15K-237-ZS candidate ranking as main.evaluate does it (main.py:230-250, utils/gen_mode_candidates.py:15-39):
    candidates = rel2candidates[rel] minus known tails minus the true tail, rank with ties//2."""
    eng = mre.engine
    z = gu.load("fb15k237_zs.npz")
    E, R, D = int(z["E"]), int(z["R"]), 200
    h, r, t = (z[k].astype(np.int64) for k in ("test_h", "test_r", "test_t"))
    ent, rel = gu.xavier_tables(gu.SEED, [(E, D), (R, D)])
    rel2cand = {int(rr): z["cand_ent"][i].astype(np.int64) for i, rr in enumerate(z["cand_rel"])}
    known = po.known_tails(h, r, t)
    # queries ordered by relation group (test_tasks iteration order already groups by relation)
    order = np.argsort(r, kind="stable")
    h, r, t = h[order], r[order], t[order]
    rels, counts = np.unique(r, return_counts=True)
    groups = eng.CandidateGroups.from_lists(counts, [rel2cand[int(x)] for x in rels], "cuda")
    # CSR of known tails per query (the e1rel_e2 filter)
    lists = [np.unique(np.asarray(known[(int(a), int(b))], np.int64)) for a, b in zip(h, r)]
    fptr = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.int64)
    fidx = np.concatenate(lists)
    rk = eng.Ranker(device=0)
    c = rk.rank("transe", (dev(ent), dev(rel)), dev(h), dev(t), dev(r), 1, groups=groups, filt_csr=(dev(fptr), dev(fidx)))
    c = c.cpu().numpy()
    ranks = c[2] + c[3] // 2 + 1
    sel = np.linspace(0, len(h) - 1, 400).astype(np.int64)
    cands = po.build_candidates(h[sel], r[sel], t[sel], rel2cand, known)
    for k, i in enumerate(sel.tolist()):
        s = ko.transe_scores(ent, rel, 1, 1, int(h[i]), int(t[i]), int(r[i]))[cands[k]]
        assert po.rank_ties_half(s) == ranks[i]
    m = rk.metrics(torch.from_numpy(c).cuda(), 1, "ties_half")
    sums, rr = m["sums"].cpu().numpy(), m["rr"].cpu().numpy()
    mrr, hits = po.summarize(ranks, (1, 3, 10))
    assert np.isclose(rr[1] / len(h), mrr, rtol=1e-12)
    assert np.isclose(sums[1][2] / len(h), hits[0]) and np.isclose(sums[1][3] / len(h), hits[1]) and np.isclose(sums[1][5] / len(h), hits[2])


def test_metrics_histogram(mre):
    eng = mre.engine
    rk = eng.Ranker(device=0)
    rng = np.random.default_rng(1)
    Q = 5000
    c = np.zeros((4, Q), np.int32)
    c[2] = rng.integers(0, 300, Q); c[3] = rng.integers(0, 4, Q); c[0] = c[2] + 5
    side = rng.integers(0, 2, Q).astype(np.uint8)
    m = rk.metrics(dev(c), dev(side), "strict", hist_len=400)
    sums, rr, hist = (m[k].cpu().numpy() for k in ("sums", "rr", "hist"))
    for s in (0, 1):
        ranks = c[2][side == s] + 1
        assert sums[s][0] == len(ranks) and sums[s][1] == ranks.sum()
        assert [sums[s][2], sums[s][3], sums[s][4], sums[s][5]] == [(ranks <= k).sum() for k in (1, 3, 5, 10)]
        assert np.isclose(rr[s], (1.0 / ranks).sum(), rtol=1e-12)
        assert sums[s][6] == int(((1 << 32) // ranks.astype(np.int64)).sum())       # 32.32 fixed-point reciprocal-rank sum
    assert np.array_equal(hist, np.bincount(c[2] + 1, minlength=400))
    raw = rk.metrics(dev(c), dev(side), "strict", raw=True)["sums"].cpu().numpy()
    assert raw[0][1] + raw[1][1] == (c[0] + 1).sum()


@pytest.mark.gpu
@pytest.mark.parametrize("scorer", ["transe", "distmult"])
def test_known_true_correction_paths_agree(mre, scorer):
    """the filtered counts do not depend on how the known-true lists reach the library: MRE_FILTER_CSR with the exact entry
    count (flattened correction pass), an over-estimate, an unknown count (one warp per query) and the same lists held by
    the index (MRE_FILTER_INDEX); repeated ids in a list count once; all equal the oracle's counts"""
    eng = mre.engine
    E, R, D = 700, 5, 72
    ds = helpers.synthetic_graph(3, E, R, 6000, 300, 500)
    rng = np.random.default_rng(4)
    ent = rng.standard_normal((E, D)).astype(np.float32)
    rel = rng.standard_normal((R, D)).astype(np.float32)
    th, tt, tr = ds.oracle.test_triples()
    side = (np.arange(len(th)) % 2).astype(np.uint8)
    ix = eng.KGIndex.from_arrays(E, R, ds.train, ds.valid, ds.test).to_device(0)
    rk = eng.Ranker(device=0)
    tabs = (dev(ent), dev(rel))
    q = (dev(th), dev(tt), dev(tr), dev(side))
    want = rk.rank(scorer, tabs, *q, index=ix).cpu().numpy()
    score = (lambda s, h, t, r: ko.transe_scores(ent, rel, 1, s, h, t, r)) if scorer == "transe" else \
            (lambda s, h, t, r: ko.distmult_scores(ent, rel, s, h, t, r))
    raw, filt = helpers.oracle_counts(ds, score, th[:60], tt[:60], tr[:60], side[:60])
    assert np.array_equal(want[0][:60], raw) and np.array_equal(want[2][:60], filt)
    # the index's lists as CSR slices (sorted; every third list carries a repeated id)
    allh, allt, allr = (np.concatenate([s[k] for s in (ds.train, ds.valid, ds.test)]) for k in range(3))
    lists = []
    for k, (h, t, r, s) in enumerate(zip(th, tt, tr, side)):
        l = np.unique(allt[(allh == h) & (allr == r)]) if s else np.unique(allh[(allt == t) & (allr == r)])
        lists.append(np.sort(np.concatenate([l, l[:1]])) if k % 3 == 0 else l)
    ptr = np.concatenate([[0], np.cumsum([len(l) for l in lists])]).astype(np.int64)
    idx = np.concatenate(lists).astype(np.int64)
    for nnz in (len(idx), len(idx) + 777, 0):
        got = rk.rank(scorer, tabs, *q, filt_csr=(dev(ptr), dev(idx), nnz)).cpu().numpy()
        assert np.array_equal(got, want), nnz
    none = rk.rank(scorer, tabs, *q).cpu().numpy()          # no lists: only the true entity leaves the filtered counts
    assert np.array_equal(none[0], want[0]) and np.array_equal(none[2], none[0]) and np.array_equal(none[3], none[1] - 1)
