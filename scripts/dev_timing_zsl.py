"""Developer timing probe: ZSL candidate scorer (entity halves + pair MLP + cosine mean + rank) at FB15K-237-ZS scale."""
import sys, os, time
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests", "golden"))
import numpy as np, torch
import mre_b200
import golden_util as gu
E, R, D, NB, T, C = 14208, 29, 200, 50, int(sys.argv[1]) if len(sys.argv) > 1 else 17596, 1000
rng = np.random.default_rng(0)
n_symbols = E + R
w = gu.seeded_extractor_weights(1, n_symbols, D)
conn = rng.integers(0, E, (E, NB)).astype(np.int64)
deg = rng.integers(1, NB + 1, E).astype(np.float32)
t0 = time.time()
ev = mre_b200.paper.ZSLEvaluator(w, conn, deg, np.arange(E), device=0)
torch.cuda.synchronize(); print(f"entity halves (E={E}): {1e3 * (time.time() - t0):.1f} ms incl. uploads")
heads = rng.integers(0, E, T); rels = rng.integers(0, R, T)
cands = [rng.choice(E, C, replace=False) for _ in range(min(T, 64))]
cands = [cands[i % len(cands)] for i in range(T)]
rel_vecs = rng.standard_normal((R, 20, D)).astype(np.float32)
ev.ctx.option("zsl_fp32", 1)
c32, s32 = ev.rank(heads, rels, cands, rel_vecs, want_scores=True)
ev.ctx.option("zsl_fp32", 0)
ctc, stc = ev.rank(heads, rels, cands, rel_vecs, want_scores=True)
d = (s32 - stc).abs()
print(f"tensor-core vs FP32 CUDA-core scores: max |diff| {d.max().item():.3e} mean {d.mean().item():.3e}; rank counts differing: {(c32[0] != ctc[0]).sum().item()} of {T}")
for it in range(2):
    counts, _ = ev.rank(heads, rels, cands, rel_vecs)
torch.cuda.synchronize()
ev.ctx.timing(True); ev.ctx.timing_read()
t0 = time.time()
counts, _ = ev.rank(heads, rels, cands, rel_vecs)
torch.cuda.synchronize(); wall = time.time() - t0
ms, n = ev.ctx.timing_read()
P = T * C
fl = 2.0 * P * (D * 2 * D)
print(f"T={T} C={C} pairs={P}: pair-MLP kernels {ms:.2f} ms ({fl / (ms * 1e-3) / 1e12:.1f} TFLOP/s algorithmic, one 400->200 contraction per pair), wall {wall * 1e3:.1f} ms incl. host list flattening, {T / wall:.4g} triples/s")
