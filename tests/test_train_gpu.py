"""GPU parity of the fused TransE margin-loss step (mre_transe_margin_step, mre_sgd_update) against the torch-CPU
restatement of the reference's TransE.forward + strategy.NegativeSampling + MarginLoss + autograd
(oracle/openke_torch.py, itself asserted bit-identical to the reference modules by tests/golden/make_golden.py).
Floating point, different summation order (and float atomics) => tolerance 1e-5 relative on loss/scores, 5e-5 of the gradient scale."""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import openke_torch as ot

pytestmark = pytest.mark.gpu


def l1_kinks(ent, rel, bh, bt, br, normalize, tol=1e-6):
    """(mask over ent gradients, mask over rel gradients) of the elements that receive a contribution from a triple whose residual
    u_d = (h^ + r^ - t^)_d is within `tol` of zero: d|u|/du jumps from -1 to +1 there, so an implementation whose row norm differs
    in the last bit may take the other one-sided derivative (each such triple moves the element by 2 / (B neg)).  Those
    elements are compared with that allowance; everything else at the stated tolerance."""
    e, r = torch.from_numpy(ent), torch.from_numpy(rel)
    h, t, rr = e[bh], e[bt], r[br]
    if normalize:
        h, t, rr = (torch.nn.functional.normalize(x, 2, -1) for x in (h, t, rr))
    near = ((h + rr) - t).abs() < tol
    me, mr = np.zeros(ent.shape, bool), np.zeros(rel.shape, bool)
    i, d = np.nonzero(near.numpy())
    me[bh[i], d] = True
    me[bt[i], d] = True
    mr[br[i], d] = True
    return me, mr


@pytest.mark.parametrize("D", [200, 300, 50])     # register-resident float4 kernels (2 and 4 groups per lane) and the scalar kernels
@pytest.mark.parametrize("p_norm,normalize,margin", [(1, True, 5.0), (2, True, 5.0), (1, False, 3.0), (2, False, 1.0)])
def test_margin_step_matches_torch_autograd(mre, fb15k237, p_norm, normalize, margin, D):
    eng = mre.engine
    E, R = fb15k237.E, fb15k237.R
    B, neg = 512, 25
    ent, rel = gu.xavier_tables(gu.SEED, [(E, D), (R, D)])
    if not normalize:
        ent, rel = ent * 30, rel * 30       # make some margins active
    bh, bt, br, by = fb15k237.oracle.sample_philox(192, 3, B, neg)
    loss_o, score_o, ge_o, gr_o = ot.transe_train_step(torch.from_numpy(ent), torch.from_numpy(rel), torch.from_numpy(bh),
                                                       torch.from_numpy(bt), torch.from_numpy(br), B, margin, p_norm, normalize)
    ctx = eng.Context(0)
    d = lambda a: torch.from_numpy(a).cuda()
    loss, ge, gr, sc = eng.transe_margin_step(ctx, d(ent), d(rel), d(bh), d(bt), d(br), B, neg, margin, p_norm, normalize,
                                              want_scores=True)
    assert np.allclose(sc.cpu().numpy(), score_o.numpy(), rtol=1e-5, atol=1e-6)
    assert np.isclose(loss.item(), loss_o.item(), rtol=1e-5)
    kinks = l1_kinks(ent, rel, bh, bt, br, normalize) if p_norm == 1 else (np.zeros(ent.shape, bool), np.zeros(rel.shape, bool))
    for mine, ref, kink in ((ge, ge_o, kinks[0]), (gr, gr_o, kinks[1])):
        ref = ref.numpy()
        scale = np.abs(ref).max()
        assert scale > 0
        err = np.abs(mine.cpu().numpy() - ref)
        assert err[~kink].max() <= 5e-5 * scale
        assert kink.mean() < 1e-2 and (err[kink].max() if kink.any() else 0.0) <= 16 * 2.0 / (B * neg)
    # SGD update (Trainer.py:73-78 with opt_method sgd): w -= lr * g, gradient buffer zeroed
    w = d(ent).clone()
    eng.sgd_update(ctx, w, ge, 0.5)
    assert np.allclose(w.cpu().numpy(), ent - 0.5 * ge_o.numpy(), rtol=1e-5, atol=1e-7)
    assert float(ge.abs().max()) == 0.0


def test_margin_step_loss_only_matches_c_oracle(mre, fb15k237):
    """loss from a given score vector: oracle/kge_oracle.c:orc_margin_loss"""
    from oracle import kge_oracle as ko
    eng = mre.engine
    E, R, D = fb15k237.E, fb15k237.R, 64
    B, neg = 100, 7
    ent, rel = gu.xavier_tables(1, [(E, D), (R, D)])
    bh, bt, br, by = fb15k237.oracle.sample_philox(5, 0, B, neg)
    ctx = eng.Context(0)
    d = lambda a: torch.from_numpy(a).cuda()
    loss, ge, gr, sc = eng.transe_margin_step(ctx, d(ent), d(rel), d(bh), d(bt), d(br), B, neg, 5.0, 1, True, want_scores=True)
    assert np.isclose(loss.item(), ko.margin_loss(sc.cpu().numpy(), B, neg, 5.0), rtol=1e-6)


def test_margin_step_config4_full_batch(mre, fb15k237):
    """BASELINE configs[3] at its own size: B = 4096 x 25 Bernoulli negatives (n = 106 496 triples), TransE L1 normalised,
    margin 5 (OpenKE/examples/train_transe_FB15K237.py:23-39) against torch autograd on the reference expressions"""
    eng = mre.engine
    E, R, D = fb15k237.E, fb15k237.R, 200
    B, neg, margin = 4096, 25, 5.0
    ent, rel = gu.xavier_tables(gu.SEED, [(E, D), (R, D)])
    bh, bt, br, by = fb15k237.oracle.sample_philox(192, 0, B, neg)
    loss_o, score_o, ge_o, gr_o = ot.transe_train_step(torch.from_numpy(ent), torch.from_numpy(rel), torch.from_numpy(bh),
                                                       torch.from_numpy(bt), torch.from_numpy(br), B, margin, 1, True)
    ctx = eng.Context(0)
    d = lambda a: torch.from_numpy(a).cuda()
    loss, ge, gr, sc = eng.transe_margin_step(ctx, d(ent), d(rel), d(bh), d(bt), d(br), B, neg, margin, 1, True, want_scores=True)
    assert np.allclose(sc.cpu().numpy(), score_o.numpy(), rtol=1e-5, atol=1e-6)
    assert np.isclose(loss.item(), loss_o.item(), rtol=1e-5)
    kinks = l1_kinks(ent, rel, bh, bt, br, True)
    for mine, ref, kink in ((ge, ge_o, kinks[0]), (gr, gr_o, kinks[1])):
        ref = ref.numpy()
        err = np.abs(mine.cpu().numpy() - ref)
        assert err[~kink].max() <= 5e-5 * np.abs(ref).max()
        assert kink.mean() < 1e-2 and (err[kink].max() if kink.any() else 0.0) <= 16 * 2.0 / (B * neg)


def _ref_scores(kind, tabs, bh, bt, br):
    if kind == "distmult":
        return ot.distmult_calc(tabs[0][bh], tabs[0][bt], tabs[1][br], "normal")
    if kind == "simple":       # OpenKE/openke/module/model/SimplE.py:19-34
        ent, rel, rel_inv = tabs
        return (torch.sum(ent[bh] * rel[br] * ent[bt], -1) + torch.sum(ent[bh] * rel_inv[br] * ent[bt], -1)) / 2
    return ot.complex_calc(tabs[0][bh], tabs[1][bh], tabs[0][bt], tabs[1][bt], tabs[2][br], tabs[3][br])


@pytest.mark.parametrize("kind", ["distmult", "complex", "simple"])
def test_similarity_models_backward_vs_autograd(mre, fb15k237, kind):
    """Model.forward of DistMult / ComplEx / SimplE must carry gradient to every embedding table
    (mre_bilinear_backward); compare .grad with torch autograd on the reference expressions"""
    ok = mre.openke
    E, R, D = fb15k237.E, fb15k237.R, 64
    B, neg = 256, 9
    bh, bt, br, by = fb15k237.oracle.sample_philox(7, 1, B, neg)
    torch.manual_seed(3)
    cls = {"distmult": ok.module.model.DistMult, "complex": ok.module.model.ComplEx, "simple": ok.module.model.SimplE}[kind]
    m = cls(E, R, dim=D).cuda()
    names = {"distmult": ("ent_embeddings", "rel_embeddings"), "simple": ("ent_embeddings", "rel_embeddings", "rel_inv_embeddings"),
             "complex": ("ent_re_embeddings", "ent_im_embeddings", "rel_re_embeddings", "rel_im_embeddings")}[kind]
    ref_tabs = [getattr(m, n).weight.detach().cpu().clone().requires_grad_() for n in names]
    w = torch.from_numpy(np.random.default_rng(2).standard_normal(B * (1 + neg)).astype(np.float32))
    data = {"batch_h": torch.from_numpy(bh).cuda(), "batch_t": torch.from_numpy(bt).cuda(), "batch_r": torch.from_numpy(br).cuda(),
            "mode": "normal"}
    score = m(data)
    assert score.requires_grad
    (score * w.cuda()).sum().backward()
    s_ref = _ref_scores(kind, ref_tabs, torch.from_numpy(bh), torch.from_numpy(bt), torch.from_numpy(br))
    (s_ref * w).sum().backward()
    assert np.allclose(score.detach().cpu().numpy(), s_ref.detach().numpy(), rtol=1e-5, atol=1e-6)
    for n, ref in zip(names, ref_tabs):
        g = getattr(m, n).weight.grad
        assert g is not None, n
        assert float((g.cpu() - ref.grad).abs().max()) <= 5e-5 * float(ref.grad.abs().max()), n


@pytest.mark.parametrize("kind,loss_name,adv", [("distmult", "SoftplusLoss", None), ("complex", "SoftplusLoss", None),
                                                ("distmult", "SigmoidLoss", 1.0), ("transe", "MarginLoss", 1.0)])
def test_strategy_recipes_fused_and_autograd(mre, fb15k237, kind, loss_name, adv):
    """the reference's own recipes (OpenKE/examples/train_distmult_WN18RR.py, train_complex_WN18RR.py: SoftplusLoss,
    regul_rate = 1.0; train_transe_FB15K237_adv-style self-adversarial losses): (a) strategy.forward + backward with the
    regulariser and (b) the fused mre_ns_train_step without it, both against CPU autograd on the reference expressions"""
    ok = mre.openke
    E, R, D = fb15k237.E, fb15k237.R, 48
    B, neg = 128, 6
    bh, bt, br, by = fb15k237.oracle.sample_philox(11, 0, B, neg)
    cls = {"distmult": ok.module.model.DistMult, "complex": ok.module.model.ComplEx, "transe": ok.module.model.TransE}[kind]
    names = {"distmult": ("ent_embeddings", "rel_embeddings"), "transe": ("ent_embeddings", "rel_embeddings"),
             "complex": ("ent_re_embeddings", "ent_im_embeddings", "rel_re_embeddings", "rel_im_embeddings")}[kind]
    lkw = dict(adv_temperature=adv) if adv is not None else {}
    if loss_name == "MarginLoss":
        lkw["margin"] = 4.0
    tb = lambda a: torch.from_numpy(a)

    def reference(tabs, regul):
        if kind == "transe":
            s = ot.transe_calc(tabs[0][tb(bh)], tabs[0][tb(bt)], tabs[1][tb(br)], "normal", 1, True)
        else:
            s = _ref_scores(kind, tabs, tb(bh), tb(bt), tb(br))
        p, n = s[:B].view(-1, B).permute(1, 0), s[B:].view(-1, B).permute(1, 0)
        T = adv if adv is not None else 0.0
        if loss_name == "SoftplusLoss":      # OpenKE/openke/module/loss/SoftplusLoss.py:22-26
            sp = torch.nn.Softplus()
            wgt = torch.softmax(n * T, -1).detach() if adv is not None else None
            val = (sp(-p).mean() + ((wgt * sp(n)).sum(-1).mean() if adv is not None else sp(n).mean())) / 2
        elif loss_name == "SigmoidLoss":     # SigmoidLoss.py:22-26
            ls = torch.nn.LogSigmoid()
            wgt = torch.softmax(n * T, -1).detach() if adv is not None else None
            val = -(ls(p).mean() + ((wgt * ls(-n)).sum(-1).mean() if adv is not None else ls(-n).mean())) / 2
        else:                                # MarginLoss.py:21-28
            mg = torch.tensor([lkw["margin"]])
            wgt = torch.softmax(-n * T, -1).detach() if adv is not None else None
            val = ((wgt * torch.max(p - n, -mg)).sum(-1).mean() if adv is not None else torch.max(p - n, -mg).mean()) + mg
        if regul:                            # DistMult.py:59-65 / ComplEx.py:42-57 / TransE.py:76-86 through strategy :27-28
            rows = [tabs[i][idx] for i, idx in ((0, tb(bh)), (0, tb(bt)), (1, tb(br)))] if kind != "complex" else \
                   [tabs[0][tb(bh)], tabs[1][tb(bh)], tabs[0][tb(bt)], tabs[1][tb(bt)], tabs[2][tb(br)], tabs[3][tb(br)]]
            val = val + regul * sum(torch.mean(x ** 2) for x in rows) / len(rows)
        return val

    for regul, fused in ((1.0, False), (0.0, True)):
        torch.manual_seed(9)
        m = cls(E, R, dim=D).cuda()
        strat = ok.module.strategy.NegativeSampling(model=m, loss=getattr(ok.module.loss, loss_name)(**lkw).cuda(), batch_size=B,
                                                    regul_rate=regul)
        ref_tabs = [getattr(m, n).weight.detach().cpu().clone().requires_grad_() for n in names]
        want = reference(ref_tabs, regul)
        want.sum().backward()
        data = {"batch_h": tb(bh).cuda(), "batch_t": tb(bt).cuda(), "batch_r": tb(br).cuda(), "batch_y": tb(by).cuda(), "mode": "normal"}
        assert strat.can_fuse() == fused
        if fused:
            loss = strat.fused_step(data)
        else:
            loss = strat(data)
            loss.backward()
        assert np.isclose(loss.item(), want.item(), rtol=2e-5), (kind, loss_name, fused)
        for n, ref in zip(names, ref_tabs):
            g = getattr(m, n).weight.grad
            assert float((g.cpu() - ref.grad).abs().max()) <= 5e-5 * float(ref.grad.abs().max()), (n, fused)
