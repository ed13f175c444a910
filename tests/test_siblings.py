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


def _rotate_scores64(ent, rel, phase_div, side, h, t, r):
    """RotatE._calc (RotatE.py:44-78) for one 1-vs-all query in float64 (the margin shift is irrelevant to the counts)"""
    Dc = rel.shape[1]
    e64, ph = ent.astype(np.float64), rel[r].astype(np.float64) / float(phase_div)
    re_r, im_r, re_e, im_e = np.cos(ph), np.sin(ph), e64[:, :Dc], e64[:, Dc:]
    if side == 0:
        re_s, im_s = (re_r * re_e[t] + im_r * im_e[t]) - re_e, (re_r * im_e[t] - im_r * re_e[t]) - im_e
    else:
        re_s, im_s = (re_e[h] * re_r - im_e[h] * im_r) - re_e, (re_e[h] * im_r + im_e[h] * re_r) - im_e
    return np.sqrt(re_s * re_s + im_s * im_s).sum(-1)


@pytest.mark.gpu
@pytest.mark.parametrize("Dc", [1, 7, 16, 33])
def test_rotate_edge_shapes(mre, Dc):
    """RotatE tile kernel on ragged shapes -- a complex dimension that is not a multiple of the 16-wide chunk (or of 4), entity and
    query counts off the 64-wide tiles, ragged candidate groups with a CSR filter, an empty job: every count inside the 1e-5 relative
    band of a float64 restatement"""
    eng = mre.engine
    rng = np.random.default_rng(40 + Dc)
    E, R, Q = 211, 5, 75
    ent = ((rng.random((E, 2 * Dc), dtype=np.float32) - 0.5) * 0.2).astype(np.float32)
    rel = ((rng.random((R, Dc), dtype=np.float32) - 0.5) * 0.2).astype(np.float32)
    div = np.float32(0.1 / np.pi)
    q_r = np.sort(rng.integers(0, R, Q))
    q_h, q_t = rng.integers(0, E, Q), rng.integers(0, E, Q)
    side = (np.arange(Q) % 2).astype(np.uint8)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    rk = eng.Ranker(device=0)
    tabs = (dev(ent), dev(rel))
    # per-query known lists (sorted), handed over as a CSR filter
    known = [np.unique(rng.integers(0, E, rng.integers(0, 6))) for _ in range(Q)]
    fptr = np.concatenate([[0], np.cumsum([len(k) for k in known])]).astype(np.int64)
    fidx = np.concatenate(known).astype(np.int64) if fptr[-1] else np.zeros(0, np.int64)
    # ragged candidate groups: one per relation present, 1 .. E candidates each
    rels, counts = np.unique(q_r, return_counts=True)
    cand = [np.sort(rng.choice(E, int(rng.integers(1, E + 1)), replace=False)) for _ in rels]
    groups = eng.CandidateGroups.from_lists(counts, cand, "cuda")
    for use_groups in (False, True):
        c = rk.rank("rotate", tabs, dev(q_h), dev(q_t), dev(q_r), dev(side), filt_csr=(dev(fptr), dev(fidx), len(fidx)),
                    groups=groups if use_groups else None, phase_div=float(div)).cpu().numpy()
        for q in range(Q):
            s = _rotate_scores64(ent, rel, div, int(side[q]), int(q_h[q]), int(q_t[q]), int(q_r[q]))
            truth = int(q_h[q]) if side[q] == 0 else int(q_t[q])
            S = np.arange(E) if not use_groups else cand[int(np.searchsorted(rels, q_r[q]))]
            band = 1e-5 * max(abs(s[truth]), float(np.abs(s).mean()))
            others = S[S != truth]
            raw_lo, raw_hi = int((s[others] < s[truth] - band).sum()), int((s[others] <= s[truth] + band).sum())
            assert raw_lo <= c[0][q] + 0 <= raw_hi and c[0][q] + c[1][q] >= raw_lo, (use_groups, q, c[:, q], raw_lo, raw_hi)
            unf = others[~np.isin(others, known[q])]
            f_lo, f_hi = int((s[unf] < s[truth] - band).sum()), int((s[unf] <= s[truth] + band).sum())
            assert f_lo <= c[2][q] <= f_hi, (use_groups, q, c[:, q], f_lo, f_hi)
    empty = torch.zeros(0, dtype=torch.int64, device="cuda")
    assert rk.rank("rotate", tabs, empty, empty, empty, 1, filter="none", phase_div=float(div)).shape == (4, 0)
