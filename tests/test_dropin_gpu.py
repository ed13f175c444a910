"""GPU: the drop-in host API (OpenKE loaders / Tester / Trainer / models, the paper's evaluate) over the C ABI gives
the reference's results: metric tuple vs the oracle's Test.h restatement, predict vs the oracle scorers, one fused
training step vs the autograd path, and main.evaluate's ranks vs the paper oracle."""
import numpy as np
import pytest
import torch

import golden_util as gu
import helpers
from oracle import kge_oracle as ko, openke_torch as ot, paper_oracle as po, ref_driver as rd

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kg(tmp_path_factory):
    ds = helpers.synthetic_graph(9, 600, 11, 9000, 400, 250)
    d = rd.write_benchmark_dir(str(tmp_path_factory.mktemp("kg")), ds.E, ds.R, ds.train, ds.valid, ds.test)
    ds.path = d
    return ds


@pytest.mark.parametrize("kind", ["transe", "distmult", "complex"])
def test_tester_tuple_matches_oracle(mre, kg, kind):
    ok = mre.openke
    torch.manual_seed(3)
    D = 48
    model = {"transe": lambda: ok.module.model.TransE(kg.E, kg.R, dim=D, p_norm=1, norm_flag=True),
             "distmult": lambda: ok.module.model.DistMult(kg.E, kg.R, dim=D),
             "complex": lambda: ok.module.model.ComplEx(kg.E, kg.R, dim=D)}[kind]()
    loader = ok.data.TestDataLoader(kg.path, "link")
    assert (loader.get_ent_tot(), loader.get_rel_tot(), loader.get_triple_tot()) == (kg.E, kg.R, kg.oracle.test_total)
    tester = ok.config.Tester(model=model, data_loader=loader, use_gpu=True)
    mrr, mr, hit10, hit3, hit1 = tester.run_link_prediction(type_constrain=False)
    tabs = [t.detach().cpu().numpy() for t in model.tables()]
    acc = ko.MetricAccumulator()
    th, tt, tr = kg.oracle.test_triples()
    for i in range(len(th)):
        for side in (0, 1):
            h, t, r = int(th[i]), int(tt[i]), int(tr[i])
            if kind == "transe":
                s = ko.transe_scores(ko.l2_normalize_rows(tabs[0]), ko.l2_normalize_rows(tabs[1]), 1, side, h, t, r)
            elif kind == "distmult":
                s = ko.distmult_scores(tabs[0], tabs[1], side, h, t, r)
            else:           # the contraction form: the association of the tcgen05 path's exact re-score -> the same counts
                s = ko.complex_scores_contracted(*tabs, side, h, t, r)
            acc.add(side, *kg.oracle.rank_from_scores(s, side, h, t, r))
    want = acc.final(kg.oracle.test_total)          # (mrr, mr, hit10, hit3, hit1) with Test.h's float32 accumulators
    assert np.allclose([mrr, hit10, hit3, hit1], [want[0], want[2], want[3], want[4]], atol=1e-4)
    assert np.isclose(mr, want[1], rtol=1e-4)
    # the per-triple loader protocol + Model.predict still work (Tester.py:77-82)
    head, tail = next(iter(loader))
    for data, side in ((head, 0), (tail, 1)):
        s = tester.test_one_step(data)
        h, t, r = int(th[0]), int(tt[0]), int(tr[0])
        ref = (ko.transe_scores(ko.l2_normalize_rows(tabs[0]), ko.l2_normalize_rows(tabs[1]), 1, side, h, t, r) if kind == "transe"
               else ko.distmult_scores(tabs[0], tabs[1], side, h, t, r) if kind == "distmult" else ko.complex_scores(*tabs, side, h, t, r))
        assert s.dtype == np.float32 and s.shape == (kg.E,)
        assert np.allclose(s, ref, rtol=2e-5, atol=1e-6)


def test_train_loader_and_fused_trainer(mre, kg):
    ok = mre.openke
    torch.manual_seed(5)
    B_total = 2048
    loader = ok.data.TrainDataLoader(in_path=kg.path, nbatches=4, threads=8, sampling_mode="normal", bern_flag=1, filter_flag=1,
                                     neg_ent=5, neg_rel=0, seed=192)
    B = loader.get_batch_size()
    assert B == kg.oracle.train_total // 4 and len(loader) == 4
    batches = list(loader)
    assert len(batches) == 4
    b0 = batches[0]
    oh, ot_, orr, oy = kg.oracle.sample_philox(192, 3, B, 5)     # the loader reuses its buffers: the last batch is step 3
    assert b0["mode"] == "normal" and b0["batch_h"].dtype == np.int64 and b0["batch_y"].dtype == np.float32
    assert np.array_equal(b0["batch_h"], oh) and np.array_equal(b0["batch_t"], ot_) and np.array_equal(b0["batch_y"], oy)

    def make():
        torch.manual_seed(7)
        m = ok.module.model.TransE(kg.E, kg.R, dim=32, p_norm=1, norm_flag=True)
        return ok.module.strategy.NegativeSampling(model=m, loss=ok.module.loss.MarginLoss(margin=5.0), batch_size=B)
    # fused SGD path vs the autograd + torch.optim.SGD path on the same Philox stream
    res = {}
    for name, opt in (("fused", "sgd"), ("autograd", "sgd")):
        strat = make()
        ld = ok.data.TrainDataLoader(in_path=kg.path, nbatches=4, bern_flag=1, neg_ent=5, seed=192, device_batches=(name == "fused"))
        tr = ok.config.Trainer(model=strat, data_loader=ld, train_times=3, alpha=0.5, use_gpu=True, opt_method=opt)
        if name == "autograd":
            strat.can_fuse = lambda: False
        tr.run()
        res[name] = (tr.losses, strat.model.ent_embeddings.weight.detach().cpu().numpy())
    assert np.allclose(res["fused"][0], res["autograd"][0], rtol=1e-4)
    assert np.allclose(res["fused"][1], res["autograd"][1], rtol=1e-3, atol=1e-5)
    assert res["fused"][0][-1] < res["fused"][0][0]          # the loss goes down
    # one step of the fused path vs the reference's tensor expressions + autograd on the CPU
    strat = make().cuda()
    ent0 = strat.model.ent_embeddings.weight.detach().cpu().clone()
    rel0 = strat.model.rel_embeddings.weight.detach().cpu().clone()
    data = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in batches[0].items()}
    loss = strat.fused_step(data)
    lo, _, ge, gr = ot.transe_train_step(ent0, rel0, data["batch_h"], data["batch_t"], data["batch_r"], B, 5.0, 1, True)
    assert np.isclose(loss.item(), lo.item(), rtol=1e-5)
    assert np.abs(strat.model.ent_embeddings.weight.grad.cpu().numpy() - ge.numpy()).max() <= 2e-5 * np.abs(ge.numpy()).max()


def test_cross_sampling_modes(mre, kg):
    ok = mre.openke
    loader = ok.data.TrainDataLoader(in_path=kg.path, batch_size=64, sampling_mode="cross", bern_flag=0, neg_ent=3, seed=1)
    it = iter(loader)
    a, b = next(it), next(it)
    assert {a["mode"], b["mode"]} == {"head_batch", "tail_batch"}
    for d in (a, b):
        if d["mode"] == "head_batch":     # heads corrupted: full batch_h, one t / r per positive (TransE.py:51-54)
            assert len(d["batch_h"]) == 64 * 4 and len(d["batch_t"]) == 64 and len(d["batch_r"]) == 64
        else:
            assert len(d["batch_t"]) == 64 * 4 and len(d["batch_h"]) == 64 and len(d["batch_r"]) == 64
    # the setters re-size the host batch arrays (the library writes B (1 + neg) elements into them)
    loader.set_ent_neg_rate(9)
    loader.set_batch_size(128)
    c = loader.cross_sampling()
    big = c["batch_h"] if c["mode"] == "head_batch" else c["batch_t"]
    assert len(big) == 128 * 10 and loader.batch_seq_size == 128 * 10
    torch.manual_seed(0)
    m = ok.module.model.TransE(kg.E, kg.R, dim=16).cuda()
    s = m(a)
    assert s.shape == (64 * 4,)
    h, t, r = [torch.from_numpy(np.asarray(a[k])) for k in ("batch_h", "batch_t", "batch_r")]
    ent, rel = m.ent_embeddings.weight.detach().cpu(), m.rel_embeddings.weight.detach().cpu()
    ref = ot.transe_calc(ent[h], ent[t], rel[r], a["mode"], 1, True)
    assert np.allclose(s.detach().cpu().numpy(), ref.numpy(), rtol=1e-5, atol=1e-6)


def test_paper_evaluate_matches_paper_oracle(mre):
    paper = mre.paper
    z = gu.load("fb15k237_zs.npz")
    E, R, D = int(z["E"]), int(z["R"]), 200
    h, r, t = (z[k].astype(np.int64) for k in ("test_h", "test_r", "test_t"))
    sel = np.sort(np.random.default_rng(0).choice(len(h), 1500, replace=False))
    h, r, t = h[sel], r[sel], t[sel]
    ent, rel = gu.xavier_tables(gu.SEED, [(E, D), (R, D)])
    rel2cand = {f"r{int(rr)}": [f"e{int(x)}" for x in z["cand_ent"][i]] for i, rr in enumerate(z["cand_rel"])}
    e2id = {f"e{i}": i for i in range(E)}
    r2id = {f"r{i}": i for i in range(R)}
    e1rel_e2 = {}
    for a, b, c in zip(h, r, t):
        e1rel_e2.setdefault(f"e{a}r{b}", []).append(f"e{c}")
    triples = [(f"e{a}", f"r{b}", f"e{c}") for a, b, c in zip(h, r, t)]
    cands = paper.build_test_candidates(triples, rel2cand, e1rel_e2)
    mrr, h1, h3, h10 = paper.evaluate(torch.from_numpy(ent), torch.from_numpy(rel), e2id, r2id, cands, verbose=False)
    ranks = []
    for rel_name, items in cands.items():
        for key, lst in items.items():
            hh = e2id[key.split("\t")[0]]
            ids = np.asarray([e2id[c] for c in lst])
            s = po.paper_transe_scores(ent, rel, hh, r2id[rel_name], ids)
            s_exact = ko.transe_scores(ent, rel, 1, 1, hh, int(ids[0]), r2id[rel_name])[ids]
            ranks.append(po.rank_ties_half(s_exact))
            assert abs(po.rank_ties_half(s) - ranks[-1]) <= 2      # numpy's summation order vs sequential: tie band only
    want_mrr, want_hits = po.summarize(ranks, (1, 3, 10))
    assert np.isclose(mrr, want_mrr, rtol=1e-12) and np.allclose([h1, h3, h10], want_hits)
    sc = paper.PaperScorer()
    hrow, rrow, trow = torch.from_numpy(ent[:50]), torch.from_numpy(rel[:50]), torch.from_numpy(ent[100:150])
    got = sc.evaluate(hrow, rrow, trow).cpu().numpy()
    assert np.allclose(got, np.abs(ent[:50] + rel[:50] - ent[100:150]).sum(-1), rtol=1e-5)
    assert paper.zsl_rank_metrics([np.array([0.9, 0.1, 0.5]), np.array([0.1, 0.9, 0.5])])[2] == (1 + 1 / 3) / 2


def test_tester_runs_rotate_through_its_own_kernel(mre, kg):
    """Tester.run_link_prediction with a RotatE model (scorer "rotate", csrc/rotate_rank.cu): the metric tuple against a numpy
    float32 restatement of RotatE._calc (OpenKE/openke/module/model/RotatE.py:44-78) ranked by the oracle's Test.h loop"""
    ok = mre.openke
    torch.manual_seed(5)
    Dc = 24
    model = ok.module.model.RotatE(kg.E, kg.R, dim=Dc, margin=6.0, epsilon=2.0)
    loader = ok.data.TestDataLoader(kg.path, "link")
    tester = ok.config.Tester(model=model, data_loader=loader, use_gpu=True)
    mrr, mr, hit10, hit3, hit1 = tester.run_link_prediction(type_constrain=False)
    ent, rel = (t.detach().cpu().numpy() for t in model.tables())
    phase = rel / np.float32(model.phase_div())
    re_r, im_r = np.cos(phase), np.sin(phase)
    re_e, im_e = ent[:, :Dc], ent[:, Dc:]
    acc = ko.MetricAccumulator()
    th, tt, tr = kg.oracle.test_triples()
    for i in range(len(th)):
        h, t, r = int(th[i]), int(tt[i]), int(tr[i])
        for side in (0, 1):
            if side == 0:        # head_batch: conj(r) o t - h
                re_s = (re_r[r] * re_e[t] + im_r[r] * im_e[t]) - re_e
                im_s = (re_r[r] * im_e[t] - im_r[r] * re_e[t]) - im_e
            else:                # tail_batch: h o r - t
                re_s = (re_e[h] * re_r[r] - im_e[h] * im_r[r]) - re_e
                im_s = (re_e[h] * im_r[r] + im_e[h] * re_r[r]) - im_e
            s = np.sqrt(re_s * re_s + im_s * im_s).sum(-1, dtype=np.float32) - np.float32(6.0)
            acc.add(side, *kg.oracle.rank_from_scores(np.ascontiguousarray(s, np.float32), side, h, t, r))
    want = acc.final(kg.oracle.test_total)
    assert np.allclose([mrr, hit10, hit3, hit1], [want[0], want[2], want[3], want[4]], atol=1e-3)
    assert np.isclose(mr, want[1], rtol=1e-3)
    head, tail = next(iter(loader))
    s = tester.test_one_step(tail)
    assert s.dtype == np.float32 and s.shape == (kg.E,)
