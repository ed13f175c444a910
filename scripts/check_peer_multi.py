"""Multi-process check of the peer-memory exchanges (run under torchrun on a multi-GPU box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_peer_multi.py
every rank: fused dp SGD step vs NCCL all-reduce + update on the same gradients (bit-identical replicas, values within FP32
re-association of the NCCL sum), the int64 all-reduce vs dist.all_reduce, and timings of both forms at the FB15K237 size."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import mre_b200

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dctx = mre_b200.dist.DistContext()
eng = mre_b200.engine
ctx = eng.Context(local)
n = (14541 + 237) * 200
pg = mre_b200.dist.PeerGroup(ctx, n)
g = torch.Generator(device="cuda").manual_seed(1)
w0 = torch.randn(n, device="cuda", generator=g)
g.manual_seed(100 + rank)
grad = torch.randn(n, device="cuda", generator=g)
pg.weights.copy_(w0)
dist.barrier(); torch.cuda.synchronize()
ref_w = w0.clone()
for step in range(3):
    pg.grads.copy_(grad * (step + 1))
    gs = grad * (step + 1)
    pg.sgd_step(0.25 / world)
    dist.all_reduce(gs)
    ref_w -= (0.25 / world) * gs
torch.cuda.synchronize()
pg.check()
err = (pg.weights - ref_w).abs().max().item()
ws = [torch.zeros_like(pg.weights) for _ in range(world)]
dist.all_gather(ws, pg.weights.clone())
same = all(torch.equal(x, ws[0]) for x in ws)
v = torch.arange(16, device="cuda", dtype=torch.int64) * (rank + 1) - 7
want = v.clone(); dist.all_reduce(want)
pg.all_reduce_i64(v)
torch.cuda.synchronize()
ok_i64 = torch.equal(v, want)

def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()

tab, gr = torch.randn(n, device="cuda"), torch.zeros(n, device="cuda")
def nccl_step():
    dist.all_reduce(gr)
    eng.sgd_update(ctx, tab, gr, 0.0)
t_fused = timeit(lambda: pg.sgd_step(0.0))
t_nccl = timeit(nccl_step)
t_i64 = timeit(lambda: pg.all_reduce_i64(v))
small = torch.zeros(16, device="cuda", dtype=torch.int64)
t_i64_nccl = timeit(lambda: dist.all_reduce(small))
pg.check()
if rank == 0:
    print(f"world {world}: max |w_fused - w_nccl| = {err:.3e} (FP32 re-association), replicas identical: {same}, int64 all-reduce ok: {ok_i64}")
    print(f"  11.8 MB gradient exchange + update: fused {t_fused * 1e3:.1f} us, NCCL all-reduce + SGD kernel {t_nccl * 1e3:.1f} us")
    print(f"  int64[16] all-reduce: peer {t_i64 * 1e3:.1f} us, NCCL {t_i64_nccl * 1e3:.1f} us")
    assert same and ok_i64 and err < 1e-4
pg.close()
dist.destroy_process_group()
