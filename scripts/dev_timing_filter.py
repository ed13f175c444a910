"""Developer probe: cost of the known-true correction on the OpenKE FB15K237 protocol."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import numpy as np, torch
import golden_util as gu
import mre_b200
eng = mre_b200.engine
z = gu.load("fb15k237_ids.npz")
E, R, D = int(z["E"]), int(z["R"]), 200
splits = tuple(gu.split_cols(z, s) for s in ("train", "valid", "test"))
ix = eng.KGIndex.from_arrays(E, R, *splits).to_device(0)
ctx = eng.Context(0); rk = eng.Ranker(ctx)
ent, rel = (torch.from_numpy(t).cuda() for t in gu.xavier_tables(192, [(E, D), (R, D)]))
th, tt, tr = ix.test_triples()
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
qh, qt, qr = d(np.repeat(th, 2)), d(np.repeat(tt, 2)), d(np.repeat(tr, 2))
side = d(np.tile(np.array([0, 1], np.uint8), len(th)))
def run(name, **kw):
    for _ in range(2): rk.rank("transe", (ent, rel), qh, qt, qr, kw.pop("side_", side), **kw) if False else None
    for _ in range(2): c = rk.rank("transe", (ent, rel), qh, qt, qr, side, **kw)
    torch.cuda.synchronize(); ctx.timing(True); ctx.timing_read()
    for _ in range(3): c = rk.rank("transe", (ent, rel), qh, qt, qr, side, **kw)
    torch.cuda.synchronize(); ms, n = ctx.timing_read(); ctx.timing(False)
    print(f"{name}: kernel {ms/n:.3f} ms")
run("no filter", filter="none")
run("index filter", index=ix)
run("index filter + normalize", index=ix, normalize=True)
# only tail queries / only head queries with the index filter
for s, nm in ((1, "tail only"), (0, "head only")):
    sel = torch.arange(s, 2 * len(th), 2, device="cuda")
    a, b, c_ = qh[sel].contiguous(), qt[sel].contiguous(), qr[sel].contiguous()
    for _ in range(2): rk.rank("transe", (ent, rel), a, b, c_, s, index=ix)
    torch.cuda.synchronize(); ctx.timing(True); ctx.timing_read()
    for _ in range(3): rk.rank("transe", (ent, rel), a, b, c_, s, index=ix)
    torch.cuda.synchronize(); ms, n = ctx.timing_read(); ctx.timing(False)
    print(f"{nm} (index filter): kernel {ms/n:.3f} ms for {len(sel)} queries")
