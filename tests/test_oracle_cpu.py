"""CPU: pin the oracle (oracle/) against the committed golden vectors produced by the REAL reference
(tests/golden/make_golden.py: the reference's OpenKE torch modules + the compiled Base.so), and against the
compiled reference library itself when oracle/_ref/Base.so is present."""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import kge_oracle as ko, openke_torch as ot, paper_oracle as po, ref_driver as rd


def test_philox_known_answers():
    assert ko.lib().orc_philox_selftest() == 0
    assert ko.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)


def test_index_tables_match_reference_stats(fb15k237):
    o = fb15k237.oracle
    assert (o.train_total, o.valid_total, o.test_total, o.triple_total) == (272115, 17535, 20466, 310116)
    g = gu.load("golden_sampler.npz")
    lm, rm = o.means()
    assert np.array_equal(lm, g["left_mean"]) and np.array_equal(rm, g["right_mean"])   # Base.so's tph / hpt, bit for bit
    th, tt, tr = o.test_triples()
    assert np.all(np.diff(tr) >= 0)


def test_lcg_sampler_reproduces_base_so_golden(fb15k237):
    """orc_sample_lcg with the seeds randReset handed out == Base.so `sampling` output (2 threads, 2 batches)"""
    g = gu.load("golden_sampler.npz")
    seeds = [int(x) for x in g["seeds"]]
    B, neg = int(g["B"]), int(g["neg"])
    for step in range(2):
        h, t, r, y = fb15k237.oracle.sample_lcg(seeds, B, neg, 0, 1)
        assert np.array_equal(h, g[f"h{step}"]) and np.array_equal(t, g[f"t{step}"]) and np.array_equal(r, g[f"r{step}"])
        assert fb15k237.oracle.count_train_leaks(h, t, r, B, B * (1 + neg)) == 0


@pytest.mark.parametrize("name,kind,kw", [
    ("transe_l1_norm", "transe", dict(p_norm=1, norm_flag=True)), ("transe_l2_norm", "transe", dict(p_norm=2, norm_flag=True)),
    ("transe_l1_raw", "transe", dict(p_norm=1, norm_flag=False)), ("distmult", "distmult", {}), ("complex", "complex", {})])
def test_rank_counts_reproduce_reference_golden(fb15k237, name, kind, kw):
    """torch restatement of Model.predict + C restatement of testHead/testTail == the reference's raw/filtered counts"""
    g = gu.load("golden_fb15k237.npz")
    E, R, D = fb15k237.E, fb15k237.R, int(g["D"])
    th, tt, tr = fb15k237.oracle.test_triples()
    ar = torch.arange(E)
    for wname, wfn in gu.WEIGHT_SETS.items():
        ent, rel, ent_im, rel_im = wfn(gu.SEED, [(E, D), (R, D), (E, D), (R, D)])
        tables = tuple(torch.from_numpy(x) for x in ((ent, ent_im, rel, rel_im) if kind == "complex" else (ent, rel)))
        for k in range(0, len(g["qidx"]), 8):
            i = int(g["qidx"][k])
            h, t, r = int(th[i]), int(tt[i]), int(tr[i])
            for side in (0, 1):
                data = ({"batch_h": ar, "batch_t": torch.tensor([t]), "batch_r": torch.tensor([r]), "mode": "head_batch"} if side == 0
                        else {"batch_h": torch.tensor([h]), "batch_t": ar, "batch_r": torch.tensor([r]), "mode": "tail_batch"})
                with torch.no_grad():
                    s = ot.predict(kind, tables, data, **kw).numpy()
                raw, filt = fb15k237.oracle.rank_from_scores(s, side, h, t, r)
                key = f"{wname}_{name}"
                assert np.allclose(s[g["probe"]], g[key + "_probe_scores"][k, side], rtol=2e-6, atol=1e-7)
                assert g[key + "_lo"][k, side] <= filt <= g[key + "_hi"][k, side]
                if np.array_equal(s[g["probe"]], g[key + "_probe_scores"][k, side]):   # same torch build => same bits
                    assert (raw, filt) == (int(g[key + "_raw"][k, side]), int(g[key + "_filt"][k, side]))


def test_c_scorers_within_tie_band_of_torch(fb15k237):
    """the fixed-order C scorers (what the CUDA kernels are bit-compared with) agree with torch's to ~1 ulp of the sum"""
    E, R, D = fb15k237.E, fb15k237.R, 200
    ent, rel, ent_im, rel_im = gu.xavier_tables(3, [(E, D), (R, D), (E, D), (R, D)])
    te, tr_ = torch.from_numpy(ent), torch.from_numpy(rel)
    ar = torch.arange(E)
    h, t, r = 11, 4242, 17
    for side, data in ((0, {"batch_h": ar, "batch_t": torch.tensor([t]), "batch_r": torch.tensor([r]), "mode": "head_batch"}),
                       (1, {"batch_h": torch.tensor([h]), "batch_t": ar, "batch_r": torch.tensor([r]), "mode": "tail_batch"})):
        for p in (1, 2):
            ref = ot.predict("transe", (te, tr_), data, p_norm=p, norm_flag=True).numpy()
            mine = ko.transe_scores(ko.l2_normalize_rows(ent), ko.l2_normalize_rows(rel), p, side, h, t, r)
            assert np.allclose(mine, ref, rtol=3e-6)
        ref = ot.predict("distmult", (te, tr_), data).numpy()
        assert np.allclose(ko.distmult_scores(ent, rel, side, h, t, r), ref, rtol=1e-4, atol=1e-9)
        ref = ot.predict("complex", tuple(torch.from_numpy(x) for x in (ent, ent_im, rel, rel_im)), data).numpy()
        assert np.allclose(ko.complex_scores(ent, ent_im, rel, rel_im, side, h, t, r), ref, rtol=1e-4, atol=1e-9)


def test_metric_accumulator_matches_reference_tuple(fb15k237):
    g = gu.load("golden_fb15k237.npz")
    acc = ko.MetricAccumulator()
    key = "structured_transe_l1_norm"
    for k in range(len(g["qidx"])):
        for side in (0, 1):
            acc.add(side, int(g[key + "_raw"][k, side]), int(g[key + "_filt"][k, side]))
    assert np.array_equal(np.asarray(acc.final(fb15k237.oracle.test_total), np.float32), g[key + "_tuple"])


def test_philox_sampler_properties(fb15k237):
    o = fb15k237.oracle
    B, neg = 2048, 8
    h, t, r, y = o.sample_philox(192, 0, B, neg)
    assert o.count_train_leaks(h, t, r, B, B * (1 + neg)) == 0
    assert o.count_train_leaks(h, t, r, 0, B) == B
    h2 = o.sample_philox(192, 1, B, neg)[0]
    assert not np.array_equal(h, h2)
    # exact-uniform skip draw: corrupt_head never returns a known train tail and covers the complement
    th, tt, tr = o.train_triples()
    hh, rr = int(th[0]), int(tr[0])
    known = set(tt[(th == hh) & (tr == rr)].tolist())
    seen = {o.corrupt_head(hh, rr, w) for w in range(0, 3 * fb15k237.E)}
    assert not (seen & known) and len(seen) == fb15k237.E - len(known)


def test_paper_oracle_rank_conventions():
    s = np.array([0.5, 0.1, 0.5, 0.5, 0.9], np.float32)
    assert po.rank_ties_half(s) == 1 + 1 + 2 // 2 == ko.rank_ties_half(s)
    assert po.rank_argsort_desc(-s) in (2, 3, 4)
    mrr, hits = po.summarize([1, 2, 4], (1, 3))
    assert np.isclose(mrr, (1 + 0.5 + 0.25) / 3) and hits == [1 / 3, 2 / 3]


@pytest.mark.skipif(not rd.available(), reason="oracle/_ref/Base.so not built (needs /root/reference)")
def test_c_restatement_equals_compiled_reference(tmp_path):
    """a small random graph through the UNMODIFIED Base.so and through kge_oracle.c: same tables, same filtered ranks,
    same metric tuple, same LCG sampler stream.  Runs in a subprocess: Base.so keeps one dataset in process globals."""
    import subprocess, sys, textwrap
    code = textwrap.dedent('''
        import sys, numpy as np
        sys.path.insert(0, %r)
        from oracle import kge_oracle as ko, ref_driver as rd
        rng = np.random.default_rng(5)
        E, R = 400, 9
        sp = lambda n: (rng.integers(0, E, n), rng.integers(0, E, n), rng.integers(0, R, n))
        tr, va, te = sp(6000), sp(300), sp(200)
        d = rd.write_benchmark_dir(%r, E, R, tr, va, te)
        ref = rd.RefOpenKE(d, threads=2)
        ix = ko.OracleIndex(E, R, tr, va, te)
        assert (ref.ent_tot, ref.rel_tot, ref.train_tot) == (E, R, ix.train_total)
        seeds = rd.RefOpenKE.lcg_seeds(2)
        for step in range(3):
            a = ref.sampling(128, 6); b = ix.sample_lcg(seeds, 128, 6)
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
        ref.load_test()
        acc = ko.MetricAccumulator()
        th, tt, trr = ix.test_triples()
        for i in range(ix.test_total):
            for side in (0, 1):
                s = rng.standard_normal(E).astype(np.float32)
                (ref.test_head if side == 0 else ref.test_tail)(s, i)
                acc.add(side, *ix.rank_from_scores(s, side, int(th[i]), int(tt[i]), int(trr[i])))
        assert ref.finish() == acc.final(ix.test_total)
        print("REF_OK")
    ''') % (str(__import__("pathlib").Path(__file__).resolve().parents[1]), str(tmp_path / "bench"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "REF_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
