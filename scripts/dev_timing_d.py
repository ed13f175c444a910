"""Developer timing probe: TransE rank kernel vs D (chunk-tail cost)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mre_b200
eng = mre_b200.engine
ctx = eng.Context(0)
rk = eng.Ranker(ctx)
peak = ctx.probe_fp32_peak()
print("fp32 peak %.4g" % peak)
E, Q = 14541, 40960
for D in [int(x) for x in (sys.argv[1:] or "128 160 192 196 200 208 224 256".split())]:
    g = torch.Generator(device="cuda").manual_seed(1)
    ent = torch.randn(E, D, device="cuda", generator=g) / D ** 0.5
    rel = torch.randn(1000, D, device="cuda", generator=g) / D ** 0.5
    qh = torch.randint(0, E, (Q,), device="cuda", generator=g)
    qt = torch.randint(0, E, (Q,), device="cuda", generator=g)
    qr = torch.randint(0, 1000, (Q,), device="cuda", generator=g)
    for it in range(2):
        c = rk.rank("transe", (ent, rel), qh, qt, qr, 1, p_norm=1)
    torch.cuda.synchronize()
    ctx.timing(True); ctx.timing_read()
    for it in range(5):
        c = rk.rank("transe", (ent, rel), qh, qt, qr, 1, p_norm=1)
    torch.cuda.synchronize()
    ms, n = ctx.timing_read()
    ctx.timing(False)
    ops = 2.0 * Q * E * D
    items = ((Q + 255) // 256) * ((E + 127) // 128)
    rounds = items / 148
    print(f"D={D}: kernel {ms/n:.3f} ms  frac {ops/(ms/n*1e-3)/peak:.3f}  us/item-round {1e3*ms/n/rounds:.2f}  ideal {256*128*D*2/(peak/148)*1e6:.2f}")
