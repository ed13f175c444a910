"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (mre_b200.dist) -- contiguous query shards,
integer metric-sum / rank-histogram all-reduce, gradient all-reduce -- gives exactly the single-process result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ranks_all, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mre_b200
    d = mre_b200.dist.DistContext()
    assert (d.rank, d.world) == (rank, world)
    lo, hi = d.shard(len(ranks_all))
    mine = ranks_all[lo:hi]
    # what mre_metrics produces for a shard: integer sums per side (all on side 1 here) + float64 rr + histogram
    sums = torch.zeros((2, 8), dtype=torch.int64)
    sums[1, 0] = len(mine); sums[1, 1] = int(mine.sum())
    for col, k in ((2, 1), (3, 3), (4, 5), (5, 10)):
        sums[1, col] = int((mine <= k).sum())
    sums[1, 6] = int(((1 << 32) // mine).sum())                   # mre_metrics slot 6: sum floor(2^32 / rank), 32.32 fixed point
    rr = torch.tensor([0.0, float((1.0 / mine).sum())], dtype=torch.float64)
    hist = torch.from_numpy(np.bincount(mine, minlength=64).astype(np.int64))
    sums2, rr2 = d.all_reduce_metrics(sums, rr)
    hist2 = d.all_reduce_hist(hist)
    g = [torch.full((4, 3), float(rank + 1)), torch.full((2,), float(10 * (rank + 1)))]
    d.all_reduce_grads(g)
    flat = torch.arange(6, dtype=torch.float32) * (rank + 1)      # the one-collective form: plain sum, 1/world folded into the SGD step
    d.all_reduce_flat(flat)
    if rank == 0:
        torch.save({"sums": sums2, "rr": rr2, "hist": hist2, "g0": g[0], "g1": g[1], "flat": flat}, os.path.join(out_dir, "out.pt"))
    d.barrier()
    dist.destroy_process_group()


def test_shards_cover_everything_once():
    sys.path.insert(0, ROOT)
    import mre_b200
    D = mre_b200.dist.DistContext
    for n in (0, 1, 7, 8, 9, 5653, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            blocks = [D.shard_of(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gloo_metrics_equal_single_process(tmp_path):
    sys.path.insert(0, ROOT)
    import mre_b200
    rng = np.random.default_rng(0)
    ranks_all = rng.integers(1, 60, 1001).astype(np.int64)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ranks_all, str(tmp_path)), nprocs=2, join=True)
    out = torch.load(os.path.join(str(tmp_path), "out.pt"))
    sums, rr, hist = out["sums"].numpy(), out["rr"].numpy(), out["hist"].numpy()
    assert sums[1][0] == len(ranks_all) and sums[1][1] == ranks_all.sum()
    assert [sums[1][2], sums[1][3], sums[1][4], sums[1][5]] == [(ranks_all <= k).sum() for k in (1, 3, 5, 10)]
    # ONE integer collective: the reciprocal-rank sum comes back from the fixed-point slot -- exactly the single-process
    # integer, hence the same MRR bits for any world size; within 2^-32 per query of the float64 sum
    assert sums[1][6] == ((1 << 32) // ranks_all).sum() and rr[1] == float(sums[1][6]) / float(1 << 32)
    assert abs(rr[1] - (1.0 / ranks_all).sum()) <= len(ranks_all) * 2.0 ** -32
    s1 = mre_b200.engine.summarize(sums)[1]
    assert s1["n"] == len(ranks_all) and abs(s1["mrr"] - (1.0 / ranks_all).mean()) <= 2.0 ** -32
    assert np.array_equal(hist, np.bincount(ranks_all, minlength=64))
    # the histogram route: float64 from integers only => identical for any world size
    m = mre_b200.dist.metrics_from_hist(hist)
    one = mre_b200.dist.metrics_from_hist(np.bincount(ranks_all, minlength=64))
    assert m == one
    assert np.isclose(m["mrr"], (1.0 / ranks_all).mean(), rtol=1e-13) and np.isclose(m["mr"], ranks_all.mean())
    assert np.isclose(m["hits10"], (ranks_all <= 10).mean()) and np.isclose(m["hits1"], (ranks_all <= 1).mean())
    # gradient all-reduce: mean over ranks
    assert torch.allclose(out["g0"], torch.full((4, 3), 1.5)) and torch.allclose(out["g1"], torch.full((2,), 15.0))
    assert torch.equal(out["flat"], torch.arange(6, dtype=torch.float32) * 3)
