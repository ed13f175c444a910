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
