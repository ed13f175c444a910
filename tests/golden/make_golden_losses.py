"""Generates tests/golden/golden_losses.npz from the REFERENCE's own loss modules (OpenKE/openke/module/loss/{MarginLoss,
SigmoidLoss,SoftplusLoss}.py; the paper's module/loss.py MarginLoss) and torch autograd: loss values AND dLoss/dp, dLoss/dn
on seeded [B, 1] / [B, neg] score blocks, plain and self-adversarial.  Build-container only; the tests read the .npz."""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/OpenKE")
import golden_util as gu  # noqa: E402

CASES = (("margin", "MarginLoss", dict(margin=5.0)), ("margin_adv", "MarginLoss", dict(adv_temperature=1.0, margin=6.0)),
         ("sigmoid", "SigmoidLoss", {}), ("sigmoid_adv", "SigmoidLoss", dict(adv_temperature=2.0)),
         ("softplus", "SoftplusLoss", {}), ("softplus_adv", "SoftplusLoss", dict(adv_temperature=0.5)))
SHAPES = ((64, 25), (257, 1), (1, 7), (300, 64))


def blocks(seed, B, neg):
    rng = np.random.default_rng([seed, B, neg])
    return (rng.standard_normal((B, 1)) * 3).astype(np.float32), (rng.standard_normal((B, neg)) * 3).astype(np.float32)


def main():
    from openke.module import loss as ref_loss
    spec = importlib.util.spec_from_file_location("paper_loss", "/root/reference/module/loss.py")
    paper_loss = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(paper_loss)
    out = {}
    for B, neg in SHAPES:
        p0, n0 = blocks(gu.SEED, B, neg)
        for name, cls, kw in CASES:
            p, n = torch.from_numpy(p0).requires_grad_(), torch.from_numpy(n0).requires_grad_()
            val = getattr(ref_loss, cls)(**kw)(p, n)
            val.sum().backward()
            tag = f"{name}_{B}x{neg}"
            out[tag + "_loss"], out[tag + "_dp"], out[tag + "_dn"] = val.detach().numpy().reshape(-1), p.grad.numpy(), n.grad.numpy()
            if cls == "MarginLoss":              # the paper's copy of the margin loss gives the same bits (module/loss.py:20-24)
                assert np.array_equal(paper_loss.MarginLoss(**kw)(torch.from_numpy(p0), torch.from_numpy(n0)).detach().numpy().reshape(-1),
                                      out[tag + "_loss"])
    np.savez_compressed(os.path.join(HERE, "golden_losses.npz"), **out)
    print("golden_losses.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
