"""Developer timing probe: the ZSL pair kernel alone (kernel ms through the context's event timer)."""
import sys, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests", "golden"))
import numpy as np, torch
import mre_b200
import golden_util as gu
E, R, D, NB, T, C = 14208, 29, 200, 50, 17596, 1000
rng = np.random.default_rng(0)
w = gu.seeded_extractor_weights(1, E + R, D)
conn = rng.integers(0, E, (E, NB)).astype(np.int64)
deg = rng.integers(1, NB + 1, E).astype(np.float32)
ev = mre_b200.paper.ZSLEvaluator(w, conn, deg, np.arange(E), device=0)
heads = rng.integers(0, E, T); rels = rng.integers(0, R, T)
cands = [rng.choice(E, C, replace=False) for _ in range(64)]
cands = [cands[i % 64] for i in range(T)]
rel_vecs = rng.standard_normal((R, 20, D)).astype(np.float32)
for it in range(2):
    ev.rank(heads, rels, cands, rel_vecs)
torch.cuda.synchronize()
ev.ctx.timing(True); ev.ctx.timing_read()
for it in range(3):
    ev.rank(heads, rels, cands, rel_vecs)
torch.cuda.synchronize()
ms, n = ev.ctx.timing_read()
print(f"{os.environ.get('MRE_B200_LIB', 'default')}: pair kernel {ms / n:.2f} ms")
