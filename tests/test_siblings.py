"""Sibling scorers and losses on the existing skeletons, against fixtures produced by the reference's own modules
(tests/golden/make_golden_siblings.py): SimplE.predict -> Base.so counts.  The losses are in tests/test_losses.py."""
import numpy as np
import pytest
import torch

import golden_util as gu


@pytest.mark.gpu
@pytest.mark.parametrize("wname", list(gu.WEIGHT_SETS))
def test_simple_ranks_vs_reference_golden(mre, fb15k237, wname):
    """SimplE ranks through the DistMult tcgen05 path: inside the reference's tie band, equal where the band is empty"""
    g = gu.load("golden_siblings.npz")
    eng = mre.engine
    E, R, D = fb15k237.E, fb15k237.R, int(g["D"])
    ent, rel, _, rel_inv = gu.WEIGHT_SETS[wname](gu.SEED, [(E, D), (R, D), (E, D), (R, D)])
    model = mre.openke.module.model.SimplE(E, R, dim=D)
    model.ent_embeddings.weight.data.copy_(torch.from_numpy(ent))
    model.rel_embeddings.weight.data.copy_(torch.from_numpy(rel))
    model.rel_inv_embeddings.weight.data.copy_(torch.from_numpy(rel_inv))
    model.cuda()
    ix = eng.KGIndex.from_arrays(E, R, fb15k237.train, fb15k237.valid, fb15k237.test).to_device(0)
    th, tt, tr = fb15k237.oracle.test_triples()
    qidx = g["qidx"]
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    q_h, q_t, q_r = np.repeat(th[qidx], 2), np.repeat(tt[qidx], 2), np.repeat(tr[qidx], 2)
    side = np.tile(np.array([0, 1], np.uint8), len(qidx))
    tabs = tuple(t.detach().contiguous() for t in model.tables())
    c = model.ranker().rank(model.scorer, tabs, dev(q_h), dev(q_t), dev(q_r), dev(side), index=ix).cpu().numpy()
    filt = c[2].reshape(-1, 2)
    lo, hi, ref = g[f"{wname}_simple_lo"], g[f"{wname}_simple_hi"], g[f"{wname}_simple_filt"]
    # head queries: the reference associates (h * r) * t, the kernel (r * t) . h -- both inside the 1e-5 relative band
    assert np.all(filt >= lo) and np.all(filt <= hi), (int(((filt < lo) | (filt > hi)).sum()),)
    exact = lo == hi
    assert np.array_equal(filt[exact], ref[exact])
    # predict parity for one query of each side
    for data in ({"batch_h": np.arange(E), "batch_t": tt[qidx[:1]], "batch_r": tr[qidx[:1]], "mode": "head_batch"},
                 {"batch_h": th[qidx[:1]], "batch_t": np.arange(E), "batch_r": tr[qidx[:1]], "mode": "tail_batch"}):
        s = model.predict({k: (torch.from_numpy(np.ascontiguousarray(v)).cuda() if k != "mode" else v) for k, v in data.items()})
        h = torch.from_numpy(ent[data["batch_h"]]); t = torch.from_numpy(ent[data["batch_t"]]); r = torch.from_numpy(rel[data["batch_r"]])
        want = -(torch.sum(h * r * t, -1)).numpy()
        assert np.allclose(s, want, rtol=2e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------------ TransH / TransD / Analogy
def build_sibling(mre, name, E, R, D, wname):
    M = mre.openke.module.model
    if name == "transh":
        m, names, shapes = M.TransH(E, R, dim=D, p_norm=1, norm_flag=True), ["ent_embeddings", "rel_embeddings", "norm_vector"], [(E, D), (R, D), (R, D)]
    elif name == "transd":
        m = M.TransD(E, R, dim_e=D, dim_r=D, p_norm=1, norm_flag=True)
        names, shapes = ["ent_embeddings", "rel_embeddings", "ent_transfer", "rel_transfer"], [(E, D), (R, D), (E, D), (R, D)]
    else:
        m = M.Analogy(E, R, dim=D)
        names = ["ent_re_embeddings", "ent_im_embeddings", "rel_re_embeddings", "rel_im_embeddings", "ent_embeddings", "rel_embeddings"]
        shapes = [(E, D), (E, D), (R, D), (R, D), (E, 2 * D), (R, 2 * D)]
    for n, t in zip(names, gu.WEIGHT_SETS[wname](gu.SEED + 7, shapes)):
        getattr(m, n).weight.data.copy_(torch.from_numpy(t))
    return m.cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("wname", list(gu.WEIGHT_SETS))
@pytest.mark.parametrize("name", ["transh", "transd", "analogy"])
def test_projected_and_analogy_ranks_vs_reference_golden(mre, fb15k237, name, wname):
    """TransH / TransD through per-relation projected tables on the TransE kernel, Analogy on the ComplEx skeleton: filtered counts
    inside the 1e-5 relative tie band of the reference module's own scores (tests/golden/golden_siblings2.npz), equal where the
    band is empty; predict() of the true entity and 16 probe entities within 2e-5 relative of the reference's scores"""
    g = gu.load("golden_siblings2.npz")
    eng = mre.engine
    E, R, D = fb15k237.E, fb15k237.R, int(g["D"])
    model = build_sibling(mre, name, E, R, D, wname)
    ix = eng.KGIndex.from_arrays(E, R, fb15k237.train, fb15k237.valid, fb15k237.test).to_device(0)
    th, tt, tr = fb15k237.oracle.test_triples()
    qidx = g["qidx"]
    q_h, q_t, q_r = np.repeat(th[qidx], 2), np.repeat(tt[qidx], 2), np.repeat(tr[qidx], 2)
    side = np.tile(np.array([0, 1], np.uint8), len(qidx))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    if hasattr(model, "rank_queries"):
        c = model.rank_queries(q_h, q_t, q_r, side, ix).cpu().numpy()
    else:
        tabs = tuple(t.detach().contiguous() for t in model.tables())
        c = model.ranker().rank(model.scorer, tabs, dev(q_h), dev(q_t), dev(q_r), dev(side), index=ix).cpu().numpy()
    key = f"{wname}_{name}"
    filt, raw = c[2].reshape(-1, 2), c[0].reshape(-1, 2)
    lo, hi, ref = g[key + "_lo"], g[key + "_hi"], g[key + "_filt"]
    outside = (filt < lo) | (filt > hi)
    assert not outside.any(), (int(outside.sum()), filt[outside][:4], lo[outside][:4], hi[outside][:4])
    exact = lo == hi
    assert np.array_equal(filt[exact], ref[exact]) and int(exact.sum()) >= 20
    assert np.all(np.abs(raw - g[key + "_raw"]) <= (hi - lo))           # the raw count moves inside the same band
    # predict(): the reference's scores of the true entity and the probe entities
    probe = g["probe"]
    for k in (0, len(qidx) // 2, len(qidx) - 1):
        i = int(qidx[k])
        h, t, r = int(th[i]), int(tt[i]), int(tr[i])
        for s in (0, 1):
            ids = np.concatenate([[h if s == 0 else t], probe])
            data = ({"batch_h": dev(ids), "batch_t": dev(np.array([t])), "batch_r": dev(np.array([r])), "mode": "head_batch"} if s == 0
                    else {"batch_h": dev(np.array([h])), "batch_t": dev(ids), "batch_r": dev(np.array([r])), "mode": "tail_batch"})
            want = g[key + "_probe"][k, s]
            assert np.allclose(model.predict(data), want, rtol=2e-5, atol=2e-6)


@pytest.mark.gpu
def test_projected_models_train_through_the_library_losses(mre):
    """TransH / TransD / Analogy forward() is differentiable to every table: one strategy step lowers the loss"""
    ok = mre.openke
    E, R, D, B, neg = 300, 7, 32, 64, 4
    rng = np.random.default_rng(0)
    h, t, r = rng.integers(0, E, B * (1 + neg)), rng.integers(0, E, B * (1 + neg)), np.tile(rng.integers(0, R, B), 1 + neg)
    data = {"batch_h": torch.from_numpy(h).cuda(), "batch_t": torch.from_numpy(t).cuda(), "batch_r": torch.from_numpy(r).cuda(),
            "batch_y": torch.ones(1).cuda(), "mode": "normal"}
    for cls, kw, loss in ((ok.module.model.TransH, dict(dim=D), ok.module.loss.MarginLoss(margin=4.0)),
                          (ok.module.model.TransD, dict(dim_e=D, dim_r=D), ok.module.loss.MarginLoss(margin=4.0)),
                          (ok.module.model.Analogy, dict(dim=D), ok.module.loss.SoftplusLoss())):
        torch.manual_seed(1)
        m = cls(E, R, **kw).cuda()
        strat = ok.module.strategy.NegativeSampling(model=m, loss=loss, batch_size=B).cuda()
        opt = torch.optim.SGD(m.parameters(), lr=0.5)
        l0 = strat(data)
        l0.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().sum() > 0 for p in m.parameters() if p.requires_grad)
        opt.step()
        with torch.no_grad():
            assert strat(data).item() < l0.item()


# ------------------------------------------------------------------------------------------------ RotatE
def build_rotate(mre, E, R, D, wname):
    m = mre.openke.module.model.RotatE(E, R, dim=D, margin=6.0, epsilon=2.0)
    ent, rel = gu.WEIGHT_SETS[wname](gu.SEED + 11, [(E, 2 * D), (R, D)])       # as tests/golden/make_golden_rotate.py
    ent = ent / np.abs(ent).max() * m.ent_embedding_range.item()
    rel = rel / np.abs(rel).max() * m.rel_embedding_range.item()
    m.ent_embeddings.weight.data.copy_(torch.from_numpy(ent.astype(np.float32)))
    m.rel_embeddings.weight.data.copy_(torch.from_numpy(rel.astype(np.float32)))
    return m.cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("wname", list(gu.WEIGHT_SETS))
def test_rotate_ranks_vs_reference_golden(mre, fb15k237, wname):
    """RotatE through its own tile kernel (csrc/rotate_rank.cu): filtered counts inside the 1e-5 relative tie band of the reference
    module's own scores (tests/golden/golden_rotate.npz), equal where the band is empty; the same counts from a CSR filter and from
    a candidate-group job over all entities; predict() within 2e-5 relative of the reference's probe scores"""
    from mre_b200.openke.module.model._projected import known_lists
    g = gu.load("golden_rotate.npz")
    eng = mre.engine
    E, R, D = fb15k237.E, fb15k237.R, int(g["D"])
    model = build_rotate(mre, E, R, D, wname)
    ix = eng.KGIndex.from_arrays(E, R, fb15k237.train, fb15k237.valid, fb15k237.test).to_device(0)
    th, tt, tr = fb15k237.oracle.test_triples()
    qidx = g["qidx"]
    q_h, q_t, q_r = np.repeat(th[qidx], 2), np.repeat(tt[qidx], 2), np.repeat(tr[qidx], 2)
    side = np.tile(np.array([0, 1], np.uint8), len(qidx))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    tabs = tuple(t.detach().contiguous() for t in model.tables())
    rk, kw = model.ranker(), model.rank_kwargs()
    c = rk.rank("rotate", tabs, dev(q_h), dev(q_t), dev(q_r), dev(side), index=ix, **kw).cpu().numpy()
    key = f"{wname}_rotate"
    filt, raw = c[2].reshape(-1, 2), c[0].reshape(-1, 2)
    lo, hi, ref = g[key + "_lo"], g[key + "_hi"], g[key + "_filt"]
    outside = (filt < lo) | (filt > hi)
    assert not outside.any(), (int(outside.sum()), filt[outside][:4], lo[outside][:4], hi[outside][:4])
    exact = lo == hi
    assert np.array_equal(filt[exact], ref[exact]) and int(exact.sum()) >= 20
    assert np.all(np.abs(raw - g[key + "_raw"]) <= (hi - lo))
    # the same job with the known lists handed over as a CSR filter, and as one candidate group holding every entity
    fptr, fidx = known_lists(ix, q_h, q_t, q_r, side)
    c2 = rk.rank("rotate", tabs, dev(q_h), dev(q_t), dev(q_r), dev(side), filt_csr=(dev(fptr), dev(fidx), len(fidx)), **kw).cpu().numpy()
    assert np.array_equal(c, c2)
    groups = eng.CandidateGroups.from_lists([len(q_h)], [np.arange(E)], "cuda")
    c3 = rk.rank("rotate", tabs, dev(q_h), dev(q_t), dev(q_r), dev(side), index=ix, groups=groups, **kw).cpu().numpy()
    assert np.array_equal(c, c3)
    probe = g["probe"]
    for k in (0, len(qidx) // 2, len(qidx) - 1):
        i = int(qidx[k])
        h, t, r = int(th[i]), int(tt[i]), int(tr[i])
        for s in (0, 1):
            ids = np.concatenate([[h if s == 0 else t], probe])
            data = ({"batch_h": dev(ids), "batch_t": dev(np.array([t])), "batch_r": dev(np.array([r])), "mode": "head_batch"} if s == 0
                    else {"batch_h": dev(np.array([h])), "batch_t": dev(ids), "batch_r": dev(np.array([r])), "mode": "tail_batch"})
            assert np.allclose(model.predict(data), g[key + "_probe"][k, s], rtol=2e-5, atol=2e-6)
    # the library's per-query score vector and tensor-core score matrix are not RotatE's: refused, not mis-dispatched
    with pytest.raises(mre._lib.MreError):
        rk.predict("rotate", tabs, dev(q_h), dev(q_t), dev(q_r), dev(side))


@pytest.mark.gpu
def test_rotate_trains_through_the_library_losses(mre):
    """RotatE.forward() is differentiable to both tables: one strategy step with the self-adversarial sigmoid loss lowers the loss"""
    ok = mre.openke
    E, R, D, B, neg = 300, 7, 16, 64, 4
    rng = np.random.default_rng(0)
    h, t, r = rng.integers(0, E, B * (1 + neg)), rng.integers(0, E, B * (1 + neg)), np.tile(rng.integers(0, R, B), 1 + neg)
    data = {"batch_h": torch.from_numpy(h).cuda(), "batch_t": torch.from_numpy(t).cuda(), "batch_r": torch.from_numpy(r).cuda(),
            "batch_y": torch.ones(1).cuda(), "mode": "normal"}
    torch.manual_seed(1)
    m = ok.module.model.RotatE(E, R, dim=D).cuda()
    strat = ok.module.strategy.NegativeSampling(model=m, loss=ok.module.loss.SigmoidLoss(adv_temperature=2), batch_size=B).cuda()
    opt = torch.optim.SGD(m.parameters(), lr=0.5)
    l0 = strat(data)
    l0.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().sum() > 0 for p in m.parameters() if p.requires_grad)
    opt.step()
    with torch.no_grad():
        assert strat(data).item() < l0.item()
