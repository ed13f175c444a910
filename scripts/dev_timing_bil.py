"""Developer timing probe: tcgen05 bilinear rank kernel."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mre_b200
eng = mre_b200.engine
ctx = eng.Context(0)
rk = eng.Ranker(ctx)
peak = ctx.probe_bf16_peak()
print("bf16 MMA peak flop/s %.4g (tf32 %.4g)" % (peak, ctx.probe_tf32_peak()))
for (kind, E, D, Q) in [("distmult", 14208, 200, 17596), ("complex", 14208, 200, 17596), ("distmult", 2_000_000, 256, 8192), ("distmult", 200_000, 256, 65536)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    mk = lambda n: torch.randn(n, D, device="cuda", generator=g) / D ** 0.5
    tabs = (mk(E), mk(1000)) if kind == "distmult" else (mk(E), mk(E), mk(1000), mk(1000))
    qh = torch.randint(0, E, (Q,), device="cuda", generator=g); qt = torch.randint(0, E, (Q,), device="cuda", generator=g)
    qr = torch.randint(0, 1000, (Q,), device="cuda", generator=g)
    for it in range(2):
        c = rk.rank(kind, tabs, qh, qt, qr, 1)
    torch.cuda.synchronize()
    ctx.timing(True); ctx.timing_read()
    t0 = time.time()
    for it in range(3):
        c = rk.rank(kind, tabs, qh, qt, qr, 1)
    torch.cuda.synchronize()
    wall = (time.time() - t0) / 3
    ms, n = ctx.timing_read(); ctx.timing(False)
    K = D * (2 if kind == "complex" else 1)
    fl = 2.0 * Q * E * K
    print(f"{kind} E={E} D={D} Q={Q}: kernel {ms/n:.3f} ms wall/step {wall*1e3:.3f} ms  {fl/(ms/n*1e-3):.4g} flop/s  frac {fl/(ms/n*1e-3)/peak:.3f} (x3 executed: {3*fl/(ms/n*1e-3)/peak:.3f})  {Q/wall:.4g} q/s  mean raw {c[0].float().mean().item():.1f}")
