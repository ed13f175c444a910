"""Developer timing probe: the RotatE tile kernel alone (kernel ms through the context's event timer), all-entity ranking."""
import os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import numpy as np, torch
import mre_b200
eng = mre_b200.engine
E, R, Dc, Q = 14541, 237, 100, 8192
rng = np.random.default_rng(0)
ent = torch.from_numpy((rng.random((E, 2 * Dc), dtype=np.float32) - 0.5) * 0.16).cuda()
rel = torch.from_numpy((rng.random((R, Dc), dtype=np.float32) - 0.5) * 0.16).cuda()
q = lambda n: torch.from_numpy(rng.integers(0, n, Q)).cuda()
q_h, q_t, q_r = q(E), q(E), q(R)
rk = eng.Ranker(device=0)
for it in range(2):
    rk.rank("rotate", (ent, rel), q_h, q_t, q_r, 1, filter="none", phase_div=0.08 / np.pi)
torch.cuda.synchronize()
rk.ctx.timing(True); rk.ctx.timing_read()
for it in range(5):
    rk.rank("rotate", (ent, rel), q_h, q_t, q_r, 1, filter="none", phase_div=0.08 / np.pi)
torch.cuda.synchronize()
ms, n = rk.ctx.timing_read()
ms /= n
elems = Q * E * Dc
peak = rk.ctx.probe_mufu_peak()
print(f"{os.environ.get('MRE_B200_LIB', 'default')}: rotate kernel {ms:.3f} ms, {elems / ms / 1e9:.2f} T complex-dim elements/s "
      f"(MUFU.SQRT probe {peak / 1e12:.2f} T/s -> {elems / ms * 1e3 / peak:.2f}; nominal 148 x 16 x 1.965 GHz = 4.65 T/s -> {elems / ms / 1e9 / 4.65:.2f})")
