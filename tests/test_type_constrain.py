"""Type-constrained link prediction (Test.h:88-98,153-163; importTypeFiles, Reader.h:267-317).

Fixture: tests/golden/golden_type_constrain.npz = FB15K237's real type_constrain.txt lists and, for 192 test triples x
2 sides x {TransE-L1-normalised, DistMult} x 2 weight sets, the constrained raw/filtered counts and the
(mrr, mr, hit10, hit3, hit1) tuple the compiled reference Base.so returned with type_constrain=1 (make_golden.py
asserts the oracle's restatement reproduces that tuple bit for bit).
CPU: the index's loader against the fixture lists and the oracle's counts against the golden ones.
GPU: grouped mre_rank (one candidate group per relation) bit-exact vs the oracle, and vs the reference golden.
"""
import numpy as np
import pytest
import torch

import golden_util as gu
import helpers
from oracle import kge_oracle as ko, ref_driver as rd


def tc_lists():
    z = gu.load("golden_type_constrain.npz")
    R = int(z["R"])
    hp, tp = z["head_ptr"], z["tail_ptr"]
    hi, ti = z["head_idx"].astype(np.int64), z["tail_idx"].astype(np.int64)
    heads = [hi[hp[r]:hp[r + 1]] for r in range(R)]
    tails = [ti[tp[r]:tp[r + 1]] for r in range(R)]
    return z, R, heads, tails


def test_index_loads_type_constrain_file(mre, fb15k237, tmp_path):
    z, R, heads, tails = tc_lists()
    eng = mre.engine
    d = rd.write_benchmark_dir(str(tmp_path / "kg"), fb15k237.E, R, fb15k237.train, fb15k237.valid, fb15k237.test)
    plain = eng.KGIndex.from_dir(d)
    assert not plain.has_type_constrain
    with pytest.raises(mre.MreError, match="no type constraints"):
        plain.type_constrain(0)
    # the file as benchmarks/*/n-n.py writes it, with the lists shuffled and one id repeated (the loader sorts + de-duplicates)
    rng = np.random.default_rng(0)
    sh = lambda l: np.concatenate([rng.permutation(l), l[:1]]) if len(l) else l
    rd.write_type_constrain(d, R, [sh(l) for l in heads], [sh(l) for l in tails])
    ix = eng.KGIndex.from_dir(d)
    assert ix.has_type_constrain
    for side, want_ptr, want_idx in ((0, z["head_ptr"], z["head_idx"]), (1, z["tail_ptr"], z["tail_idx"])):
        ptr, idx = ix.type_constrain(side)
        assert np.array_equal(ptr, want_ptr) and np.array_equal(idx, want_idx.astype(np.int64))
    # in-memory installation gives the same tables; an explicit path works too
    plain.set_type_constrain(z["head_ptr"], z["head_idx"], z["tail_ptr"], z["tail_idx"])
    assert np.array_equal(plain.type_constrain(1)[1], ix.type_constrain(1)[1])
    other = eng.KGIndex.from_arrays(fb15k237.E, R, fb15k237.train)
    other.load_type_constrain(d + "/type_constrain.txt")
    assert np.array_equal(other.type_constrain(0)[0], z["head_ptr"])
    with pytest.raises(mre.MreError, match="out of range"):
        eng.KGIndex.from_arrays(10, R, (np.array([1]), np.array([2]), np.array([0]))).load_type_constrain(d + "/type_constrain.txt")


def test_oracle_constrained_counts_match_reference_golden(fb15k237):
    """oracle/kge_oracle.c's constrained walk on the oracle's own scores lands inside the reference's counts' tie band:
    DistMult (no near-ties on these tables) must match the golden counts exactly"""
    z, R, heads, tails = tc_lists()
    E, D = fb15k237.E, 200
    th, tt, tr = fb15k237.oracle.test_triples()
    qidx = z["qidx"]
    ent, rel, _, _ = gu.WEIGHT_SETS["xavier"](gu.SEED, [(E, D), (R, D), (E, D), (R, D)])
    want_raw, want_filt = z["xavier_distmult_raw"], z["xavier_distmult_filt"]
    for k, i in enumerate(qidx[:48].tolist()):
        h, t, r = int(th[i]), int(tt[i]), int(tr[i])
        for side in (0, 1):
            s = ko.distmult_scores(ent, rel, side, h, t, r)
            raw, filt = fb15k237.oracle.rank_from_scores_constrained(s, side, h, t, r, heads[r] if side == 0 else tails[r])
            assert (raw, filt) == (want_raw[k, side], want_filt[k, side])


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def env(mre, fb15k237):
    eng = mre.engine
    z, R, heads, tails = tc_lists()
    ix = eng.KGIndex.from_arrays(fb15k237.E, fb15k237.R, fb15k237.train, fb15k237.valid, fb15k237.test)
    ix.set_type_constrain(z["head_ptr"], z["head_idx"], z["tail_ptr"], z["tail_idx"])
    ix.to_device(0)
    return eng, ix, eng.Ranker(device=0)


def groups_for(eng, ix, side, q_r):
    ptr, idx = ix.type_constrain(side)
    qptr = np.searchsorted(q_r, np.arange(ix.rel_tot + 1), side="left")
    return eng.CandidateGroups(qptr, ptr, dev(idx))


@pytest.mark.gpu
@pytest.mark.parametrize("wname", list(gu.WEIGHT_SETS))
@pytest.mark.parametrize("name", ["transe_l1_norm", "distmult"])
def test_type_constrained_counts_gpu(env, fb15k237, name, wname):
    eng, ix, rk = env
    z, R, heads, tails = tc_lists()
    E, D = fb15k237.E, 200
    ent, rel, _, _ = gu.WEIGHT_SETS[wname](gu.SEED, [(E, D), (R, D), (E, D), (R, D)])
    th, tt, tr = fb15k237.oracle.test_triples()
    qidx = z["qidx"]
    q_h, q_t, q_r = th[qidx], tt[qidx], tr[qidx]
    assert np.all(np.diff(q_r) >= 0)
    if name == "distmult":
        scorer, kw = "distmult", {}
        score = lambda s, h, t, r: ko.distmult_scores(ent, rel, s, h, t, r)
    else:
        scorer, kw = "transe", dict(p_norm=1, normalize=True)
        eo, ro = ko.l2_normalize_rows(ent), ko.l2_normalize_rows(rel)
        score = lambda s, h, t, r: ko.transe_scores(eo, ro, 1, s, h, t, r)
    sums_all = np.zeros((2, 8), np.int64)
    rr_all = np.zeros(2)
    bands = gu.load("golden_bands.npz")
    for side in (0, 1):
        c = rk.rank(scorer, (dev(ent), dev(rel)), dev(q_h), dev(q_t), dev(q_r), side, index=ix,
                    groups=groups_for(eng, ix, side, q_r), **kw)
        cn = c.cpu().numpy()
        # bit-exact against the oracle's constrained walk over the oracle's scores
        for k in range(0, len(qidx), 3):
            h, t, r = int(q_h[k]), int(q_t[k]), int(q_r[k])
            raw, filt = fb15k237.oracle.rank_from_scores_constrained(score(side, h, t, r), side, h, t, r,
                                                                      heads[r] if side == 0 else tails[r])
            assert (cn[0][k], cn[2][k]) == (raw, filt), (side, k)
        # against the compiled reference's counts: inside the 1e-5 relative tie band of the reference's own scores
        # (tests/golden/golden_bands.npz), equal wherever that band is empty
        ref_filt = z[f"{wname}_{name}_filt"][:, side]
        lo, hi = bands[f"tc_{wname}_{name}_lo"][:, side], bands[f"tc_{wname}_{name}_hi"][:, side]
        assert np.all((cn[2] >= lo) & (cn[2] <= hi)), (side, int(((cn[2] < lo) | (cn[2] > hi)).sum()))
        assert np.array_equal(cn[2][lo == hi], ref_filt[lo == hi])
        m = rk.metrics(c, side, "strict")
        sums_all += m["sums"].cpu().numpy()
        rr_all += m["rr"].cpu().numpy()
    T = float(fb15k237.oracle.test_total)
    mine = np.array([rr_all.sum() / 2 / T, sums_all[:, 1].sum() / 2 / T, sums_all[:, 5].sum() / 2 / T,
                     sums_all[:, 3].sum() / 2 / T, sums_all[:, 2].sum() / 2 / T])
    ref_t = z[f"{wname}_{name}_tuple"].astype(np.float64)
    assert np.allclose(mine[[0, 2, 3, 4]], ref_t[[0, 2, 3, 4]], atol=1e-4)
    assert np.isclose(mine[1], ref_t[1], rtol=2e-3)


@pytest.mark.gpu
def test_empty_type_lists_are_skipped(mre):
    """relations whose list is empty rank against nothing: counts 0 (rank 1), neighbours unaffected"""
    eng = mre.engine
    E, R, D = 300, 4, 8
    ds = helpers.synthetic_graph(5, E, R, 2000, 100, 400)
    ix = eng.KGIndex.from_arrays(E, R, ds.train, ds.valid, ds.test)
    rng = np.random.default_rng(2)
    lists = [np.zeros(0, np.int64), np.sort(rng.choice(E, 150, replace=False)), np.zeros(0, np.int64), np.arange(E)]
    ptr = np.concatenate([[0], np.cumsum([len(l) for l in lists])])
    idx = np.concatenate(lists)
    ix.set_type_constrain(ptr, idx, ptr, idx)
    ix.to_device(0)
    rk = eng.Ranker(device=0)
    ent = rng.standard_normal((E, D)).astype(np.float32)
    rel = rng.standard_normal((R, D)).astype(np.float32)
    th, tt, tr = ds.oracle.test_triples()
    for side in (0, 1):
        c = rk.rank("transe", (dev(ent), dev(rel)), dev(th), dev(tt), dev(tr), side, index=ix,
                    groups=groups_for(eng, ix, side, tr)).cpu().numpy()
        for k in range(len(th)):
            h, t, r = int(th[k]), int(tt[k]), int(tr[k])
            raw, filt = ds.oracle.rank_from_scores_constrained(ko.transe_scores(ent, rel, 1, side, h, t, r), side, h, t, r, lists[r])
            assert (c[0][k], c[2][k]) == (raw, filt)


# ------------------------------------------------------------------------------------------------ corrupt(h, r)
def test_oracle_corrupt_matches_compiled_reference(fb15k237):
    """oracle/kge_oracle.c's corrupt() fed the libc rand() values the compiled reference consumed (and thread 0's LCG for the
    corrupt_head fallback) returns the reference's own tails: tests/golden/golden_corrupt.npz (make_golden_corrupt.py), 2 048
    (h, r) pairs of FB15K237, 44 of them through the 1000-redraw fallback (Corrupt.h:179-195)"""
    z, R, heads, tails = tc_lists()
    g = gu.load("golden_corrupt.npz")
    h, r = g["h"].astype(np.int64), g["r"].astype(np.int64)
    got, used, _ = fb15k237.oracle.corrupt_typed_words(z["tail_ptr"], z["tail_idx"].astype(np.int64), h, r, g["words"].astype(np.int64), int(g["lcg0"]))
    assert used == len(g["words"])
    assert np.array_equal(got, g["tails"].astype(np.int64))


def corrupt_cases(fb15k237):
    """FB15K237 pairs of the reference fixture (fallbacks included) on the real tail-type lists"""
    z, R, heads, tails = tc_lists()
    g = gu.load("golden_corrupt.npz")
    return z, tails, g["h"].astype(np.int64), g["r"].astype(np.int64)


def check_corrupt_properties(orc, tails, h, r, out, all_known_falls_back=True):
    for a, b, t in zip(h.tolist(), r.tolist(), out.tolist()):
        assert not orc.find(a, t, b)                                       # never a known triple of any split ...
        exhausted = all(orc.find(a, int(x), b) for x in tails[b])
        assert exhausted or t in tails[b]                                  # ... from the type list unless it is exhausted


def test_corrupt_philox_replay_properties(fb15k237):
    z, tails, h, r = corrupt_cases(fb15k237)
    out = fb15k237.oracle.corrupt_typed_philox(192, 3, z["tail_ptr"], z["tail_idx"].astype(np.int64), h, r)
    check_corrupt_properties(fb15k237.oracle, tails, h, r, out)
    again = fb15k237.oracle.corrupt_typed_philox(192, 3, z["tail_ptr"], z["tail_idx"].astype(np.int64), h, r)
    other = fb15k237.oracle.corrupt_typed_philox(192, 4, z["tail_ptr"], z["tail_idx"].astype(np.int64), h, r)
    assert np.array_equal(out, again) and not np.array_equal(out, other)
    # uniform over the admissible part of the list: one pair, many steps
    k = int(np.argmax([len(tails[b]) for b in r]))
    adm = np.array([x for x in tails[r[k]] if not fb15k237.oracle.find(int(h[k]), int(x), int(r[k]))])
    n = 4000
    draws = np.concatenate([fb15k237.oracle.corrupt_typed_philox(5, s, z["tail_ptr"], z["tail_idx"].astype(np.int64), np.full(200, h[k]), np.full(200, r[k]))
                            for s in range(n // 200)])
    assert set(draws.tolist()) <= set(adm.tolist())
    assert len(np.unique(draws)) > 0.6 * min(len(adm), n)


@pytest.mark.gpu
def test_corrupt_typed_gpu_bit_exact(env, fb15k237, mre):
    """mre_corrupt_typed == the CPU replay of the same Philox stream, element for element (fallback pairs included)"""
    eng, ix, rk = env
    z, tails, h, r = corrupt_cases(fb15k237)
    smp = eng.Sampler(ix, ctx=rk.ctx, seed=192)
    for step in (0, 3, 1 << 33):
        out = smp.corrupt_typed(step, dev(h), dev(r)).cpu().numpy()
        want = fb15k237.oracle.corrupt_typed_philox(192, step, z["tail_ptr"], z["tail_idx"].astype(np.int64), h, r)
        assert np.array_equal(out, want)
    check_corrupt_properties(fb15k237.oracle, tails, h, r, out)
    plain = eng.KGIndex.from_arrays(fb15k237.E, fb15k237.R, fb15k237.train).to_device(0)
    with pytest.raises(mre.MreError, match="no type constraints"):
        eng.Sampler(plain, ctx=rk.ctx).corrupt_typed(0, dev(h), dev(r))


@pytest.mark.gpu
def test_corrupt_typed_gpu_empty_and_exhausted_lists(mre):
    """an empty list and a list whose every member is known both take the corrupt_head fallback; type lists installed AFTER
    the index went to the device are uploaded too"""
    eng = mre.engine
    E, R = 50, 3
    ds = helpers.synthetic_graph(9, E, R, 600, 30, 30)
    ix = eng.KGIndex.from_arrays(E, R, ds.train, ds.valid, ds.test).to_device(0)
    h0, r0 = int(ds.train[0][0]), 1
    known = np.array(sorted({int(t) for hh, t, rr in zip(*ds.train) if hh == h0 and rr == r0} | {0}))
    ds2 = helpers.synthetic_graph(9, E, R, 600, 30, 30)
    lists = [np.zeros(0, np.int64), known[known != 0] if len(known) > 1 else np.zeros(0, np.int64), np.arange(E)]
    ptr = np.concatenate([[0], np.cumsum([len(l) for l in lists])]).astype(np.int64)
    idx = np.concatenate(lists).astype(np.int64)
    ix.set_type_constrain(ptr, idx, ptr, idx)
    rng = np.random.default_rng(1)
    h = np.concatenate([[h0] * 8, rng.integers(0, E, 500)]).astype(np.int64)
    r = np.concatenate([[r0] * 8, rng.integers(0, R, 500)]).astype(np.int64)
    smp = eng.Sampler(ix, seed=7)
    out = smp.corrupt_typed(2, dev(h), dev(r)).cpu().numpy()
    want = ds2.oracle.corrupt_typed_philox(7, 2, ptr, idx, h, r)
    assert np.array_equal(out, want)
    for a, b, t in zip(h.tolist(), r.tolist(), out.tolist()):
        assert not any((hh, tt, rr) == (a, t, b) for hh, tt, rr in zip(*ds.train))     # the fallback never emits a train triple
