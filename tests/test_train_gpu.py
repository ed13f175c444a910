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


@pytest.mark.parametrize("p_norm,normalize,margin", [(1, True, 5.0), (2, True, 5.0), (1, False, 3.0), (2, False, 1.0)])
def test_margin_step_matches_torch_autograd(mre, fb15k237, p_norm, normalize, margin):
    eng = mre.engine
    E, R, D = fb15k237.E, fb15k237.R, 200
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
    for mine, ref in ((ge, ge_o), (gr, gr_o)):
        ref = ref.numpy()
        scale = np.abs(ref).max()
        assert scale > 0
        assert np.abs(mine.cpu().numpy() - ref).max() <= 5e-5 * scale
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
