"""Profiling driver: a few launches of the fused TransE rank kernel at one shape (run under ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mre_b200
E, D, Q, p = [int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (12741, 200, 5653, 1))]
eng = mre_b200.engine
rk = eng.Ranker(device=0)
g = torch.Generator(device="cuda").manual_seed(1)
ent = torch.randn(E, D, device="cuda", generator=g) / D ** 0.5
rel = torch.randn(1000, D, device="cuda", generator=g) / D ** 0.5
qh = torch.randint(0, E, (Q,), device="cuda", generator=g)
qt = torch.randint(0, E, (Q,), device="cuda", generator=g)
qr = torch.randint(0, 1000, (Q,), device="cuda", generator=g)
for it in range(3):
    c = rk.rank("transe", (ent, rel), qh, qt, qr, 1, p_norm=p)
torch.cuda.synchronize()
print("ok", c[0].float().mean().item())
