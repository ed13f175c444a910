"""Developer timing probe (not the bench contract): fused TransE rank kernel on synthetic tables."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mre_b200
eng = mre_b200.engine
ctx = eng.Context(0)
rk = eng.Ranker(ctx)
print("sm_count", ctx.sm_count)
peak = ctx.probe_fp32_peak()
print("fp32 FADD peak lane-ops/s %.4g" % peak)
SHAPES = [(12741, 200, 5653, 1), (14541, 200, 40932, 1), (2_000_000, 256, 8192, 1), (2_000_000, 256, 8192, 2), (200_000, 256, 65536, 1)]
if os.environ.get('MRE_QUICK'):
    SHAPES = [(12741, 200, 5653, 1), (200_000, 256, 16384, 1), (200_000, 256, 16384, 2)]
for (E, D, Q, p) in SHAPES:
    g = torch.Generator(device="cuda").manual_seed(1)
    ent = torch.randn(E, D, device="cuda", generator=g) / D ** 0.5
    rel = torch.randn(1000, D, device="cuda", generator=g) / D ** 0.5
    qh = torch.randint(0, E, (Q,), device="cuda", generator=g)
    qt = torch.randint(0, E, (Q,), device="cuda", generator=g)
    qr = torch.randint(0, 1000, (Q,), device="cuda", generator=g)
    for it in range(2):
        c = rk.rank("transe", (ent, rel), qh, qt, qr, 1, p_norm=p)
    torch.cuda.synchronize()
    ctx.timing(True)
    t0 = time.time()
    for it in range(3):
        c = rk.rank("transe", (ent, rel), qh, qt, qr, 1, p_norm=p)
    torch.cuda.synchronize()
    wall = (time.time() - t0) / 3
    ms, n = ctx.timing_read()
    ctx.timing(False)
    ops = 2.0 * Q * E * D
    print(f"E={E} D={D} Q={Q} p={p}: kernel {ms/n:.3f} ms  wall/step {wall*1e3:.3f} ms  {ops/(ms/n*1e-3):.4g} lane-op/s  frac {ops/(ms/n*1e-3)/peak:.3f}  {Q/(wall):.4g} q/s  mean raw {c[0].float().mean().item():.1f}")
