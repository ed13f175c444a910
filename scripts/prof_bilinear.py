"""Profiling driver: a few launches of the tcgen05 bilinear rank kernel at one shape (run under ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mre_b200
kind = sys.argv[1] if len(sys.argv) > 1 else "distmult"
E, D, Q = [int(x) for x in (sys.argv[2:5] if len(sys.argv) > 4 else (14208, 200, 17596))]
eng = mre_b200.engine
rk = eng.Ranker(device=0)
g = torch.Generator(device="cuda").manual_seed(1)
mk = lambda n: torch.randn(n, D, device="cuda", generator=g) / D ** 0.5
tabs = (mk(E), mk(1000)) if kind == "distmult" else (mk(E), mk(E), mk(1000), mk(1000))
qh = torch.randint(0, E, (Q,), device="cuda", generator=g); qt = torch.randint(0, E, (Q,), device="cuda", generator=g)
qr = torch.randint(0, 1000, (Q,), device="cuda", generator=g)
for it in range(3):
    c = rk.rank(kind, tabs, qh, qt, qr, 1)
torch.cuda.synchronize()
print("ok", c[0].float().mean().item())
