"""The negative-sampling losses (mre_ns_loss behind openke.module.loss.{MarginLoss,SigmoidLoss,SoftplusLoss}) against the
reference's own modules + torch autograd (tests/golden/golden_losses.npz, golden_siblings.npz): value within 1e-5 relative,
dLoss/dscore within 1e-5 of the gradient scale; plain and self-adversarial; ragged shapes (B = 1, neg = 1)."""
import numpy as np
import pytest
import torch

import golden_util as gu
from make_golden_losses import CASES, SHAPES, blocks


def test_losses_fail_loudly_without_gpu(mre):
    """no CPU fallback: a CPU score block raises instead of being evaluated with torch ops"""
    loss = mre.openke.module.loss
    p, n = torch.zeros(4, 1), torch.zeros(4, 3)
    for cls in (loss.MarginLoss, loss.SigmoidLoss, loss.SoftplusLoss):
        with pytest.raises(mre.MreError):
            cls()(p, n)


def test_loss_state_dict_names(mre):
    """the frozen hyper-parameters keep the reference's state_dict names (MarginLoss.py:12-18, BaseModule.py:7-12)"""
    loss = mre.openke.module.loss
    assert set(loss.MarginLoss(adv_temperature=1.0, margin=4.0).state_dict()) == {"zero_const", "pi_const", "margin", "adv_temperature"}
    assert set(loss.SoftplusLoss().state_dict()) == {"zero_const", "pi_const"}
    m = loss.MarginLoss(margin=4.0)
    assert m.margin.item() == 4.0 and not m.margin.requires_grad and not m.adv_flag


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES)
def test_losses_value_and_gradient_vs_reference(mre, shape):
    g = gu.load("golden_losses.npz")
    loss = mre.openke.module.loss
    B, neg = shape
    p0, n0 = blocks(gu.SEED, B, neg)
    for name, cls, kw in CASES:
        p = torch.from_numpy(p0).cuda().requires_grad_()
        n = torch.from_numpy(n0).cuda().requires_grad_()
        val = getattr(loss, cls)(**kw)(p, n)
        assert val.shape == (1,)
        val.sum().backward()
        tag = f"{name}_{B}x{neg}"
        assert np.allclose(val.item(), g[tag + "_loss"][0], rtol=1e-5, atol=1e-6), (tag, val.item(), g[tag + "_loss"])
        for got, want in ((p.grad, g[tag + "_dp"]), (n.grad, g[tag + "_dn"])):
            scale = max(float(np.abs(want).max()), 1e-12)
            assert np.abs(got.cpu().numpy() - want).max() <= 1e-5 * scale, tag


@pytest.mark.gpu
def test_losses_on_strategy_views(mre):
    """the strategy hands PERMUTED VIEWS of the flat score vector (strategy/NegativeSampling.py:13-21): same value, and the
    gradient lands on the flat vector in its own layout"""
    g = gu.load("golden_siblings.npz")
    loss = mre.openke.module.loss
    p0, n0 = g["loss_p"], g["loss_n"]
    B, neg = n0.shape
    flat = torch.from_numpy(np.concatenate([p0.reshape(-1), n0.T.reshape(-1)])).cuda().requires_grad_()
    p = flat[:B].view(-1, B).permute(1, 0)
    n = flat[B:].view(-1, B).permute(1, 0)
    for name, cls, kw in CASES:
        flat.grad = None
        val = getattr(loss, cls)(**kw)(p, n)
        assert np.allclose(val.item(), g["loss_" + name][0], rtol=1e-5, atol=1e-6), name
        val.backward()
        pc = torch.from_numpy(p0).cuda().requires_grad_()
        nc = torch.from_numpy(n0).cuda().requires_grad_()
        getattr(loss, cls)(**kw)(pc, nc).backward()
        assert torch.equal(flat.grad[:B], pc.grad.reshape(-1)) and torch.equal(flat.grad[B:].view(neg, B).t(), nc.grad)
