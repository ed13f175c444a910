"""The fused gradient reduction + SGD step and the int64 metric all-reduce over peer memory (csrc/peer.cu), on ONE GPU: the
"ranks" are peer groups of this process whose kernels run side by side on separate streams (every flag wait is a real
cross-kernel wait).  The multi-process form (CUDA IPC handles through torch.distributed) is exercised by
`bench.py --workload train --gpus N` and scripts/check_peer_multi.py on a multi-GPU box."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [1, 2, 3])
def test_dp_sgd_step_equals_sum_then_update(mre, world):
    eng, D = mre.engine, mre.dist
    n = 4 * 25_003                                       # float4 count not divisible by the world sizes: ragged slices
    ctxs = [eng.Context(0) for _ in range(world)]
    groups = [D.PeerGroup(c, n, rank=r, world=world, local=True) for r, c in enumerate(ctxs)]
    D.PeerGroup.connect_local(groups)
    g = torch.Generator(device="cuda").manual_seed(5)
    w0 = torch.randn(n, device="cuda", generator=g)
    grads = [torch.randn(n, device="cuda", generator=g) for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    want = w0.clone()
    for step in range(3):
        for r, pg in enumerate(groups):
            if step == 0:
                pg.weights.copy_(w0)
            pg.grads.copy_(grads[r] * (step + 1))
        torch.cuda.synchronize()
        for r, pg in enumerate(groups):                  # "ranks" launch in reverse order: early ones really wait for late ones
            with torch.cuda.stream(streams[world - 1 - r]):
                groups[world - 1 - r].sgd_step(0.5 / world, max_blocks=4)
        torch.cuda.synchronize()
        s = grads[0] * (step + 1)
        for r in range(1, world):
            s = s + grads[r] * (step + 1)                # rank order, as the kernel sums
        want = want - (0.5 / world) * s
        for pg in groups:
            pg.check()
            assert torch.equal(pg.weights, want)         # bit-identical on every rank
            assert not pg.grads.any()                    # re-armed for the next backward
    for pg in groups:
        pg.close()


def test_peer_allreduce_i64(mre):
    eng, D = mre.engine, mre.dist
    world = 4
    ctxs = [eng.Context(0) for _ in range(world)]
    groups = [D.PeerGroup(c, 64, rank=r, world=world, local=True) for r, c in enumerate(ctxs)]
    D.PeerGroup.connect_local(groups)
    streams = [torch.cuda.Stream() for _ in range(world)]
    rng = np.random.default_rng(0)
    for it in range(3):
        vals = rng.integers(-2 ** 40, 2 ** 40, (world, 16))
        vecs = [torch.from_numpy(vals[r]).cuda() for r in range(world)]
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                groups[r].all_reduce_i64(vecs[r])
        torch.cuda.synchronize()
        for r in range(world):
            groups[r].check()
            assert np.array_equal(vecs[r].cpu().numpy(), vals.sum(0))
    for pg in groups:
        pg.close()
