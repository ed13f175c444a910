"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch on the
B200 box; gloo in the CPU tests).  The path shards by QUERY: every rank ranks a contiguous block of queries against
the replicated entity table, so there is no data-path collective; the only exchanges are
  * one all-reduce of the integer metric sums / the integer rank histogram (eval), and
  * one all-reduce of the two dense gradient tables per step (data-parallel training).
MRR is computed from the all-reduced INTEGER rank histogram in float64, so it is bit-identical for any number of GPUs.
"""
import os

import numpy as np
import torch
import torch.distributed as dist


class DistContext:
    def __init__(self, backend=None, device=None):
        if not dist.is_initialized():
            backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            kw = {}
            if backend == "nccl":
                local = int(os.environ.get("LOCAL_RANK", "0"))
                torch.cuda.set_device(local)
                kw["device_id"] = torch.device("cuda", local)
            dist.init_process_group(backend, **kw)
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.device = device

    def shard(self, n):
        """contiguous block [lo, hi) of n units owned by this rank (blocks differ in size by at most one)"""
        base, rem = divmod(n, self.world)
        lo = self.rank * base + min(self.rank, rem)
        return lo, lo + base + (1 if self.rank < rem else 0)

    @staticmethod
    def shard_of(n, rank, world):
        base, rem = divmod(n, world)
        lo = rank * base + min(rank, rem)
        return lo, lo + base + (1 if rank < rem else 0)

    def all_reduce_metrics(self, sums, rr):
        """sums: int64 [2, 8]; rr: float64 [2] -> summed over ranks (in place on copies)"""
        sums, rr = sums.clone(), rr.clone()
        dist.all_reduce(sums)
        dist.all_reduce(rr)
        return sums, rr

    def all_reduce_hist(self, hist):
        hist = hist.clone()
        dist.all_reduce(hist)
        return hist

    def all_reduce_grads(self, grads):
        """sum each gradient table over ranks and scale by 1/world (data-parallel mean of per-rank mean losses)"""
        for g in grads:
            dist.all_reduce(g)
            g.mul_(1.0 / self.world)

    def all_reduce_flat(self, flat):
        """ONE sum all-reduce of a flat gradient buffer holding every table back to back; no scaling here -- the caller folds
        1/world into its SGD step (w -= (lr / world) * sum g), so a data-parallel step is one collective + one update kernel"""
        dist.all_reduce(flat)
        return flat

    def barrier(self):
        dist.barrier()


def metrics_from_hist(hist):
    """rank histogram (hist[k] = #queries with rank k) -> dict(n, mr, mrr, hits1/3/5/10); float64 from integers only"""
    h = np.asarray(hist.cpu() if isinstance(hist, torch.Tensor) else hist, dtype=np.int64)
    k = np.arange(len(h), dtype=np.float64)
    n = int(h.sum())
    if n == 0:
        return None
    nz = np.nonzero(h)[0]
    nz = nz[nz > 0]
    mrr = float(np.sum(h[nz] / k[nz])) / n
    mr = float(np.sum(h[nz] * k[nz])) / n
    cum = np.cumsum(h)
    hit = lambda x: float(cum[min(x, len(h) - 1)]) / n
    return {"n": n, "mr": mr, "mrr": mrr, "hits1": hit(1), "hits3": hit(3), "hits5": hit(5), "hits10": hit(10)}
