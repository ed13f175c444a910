"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch on the
B200 box; gloo in the CPU tests).  The path shards by QUERY: every rank ranks a contiguous block of queries against
the replicated entity table, so there is no data-path collective; the only exchanges are
  * ONE all-reduce of the int64 [2, 8] metric sums (eval) -- counts, rank sums, Hits@k and the reciprocal-rank sum in 32.32
    fixed point (mre_metrics slot 6), all integers, so the reduced tuple is bit-identical for any number of GPUs and any
    sharding (MRR to within 2^-32 of the float64 sum); the integer rank histogram is the alternative exact form
    (metrics_from_hist; it needs hist_len > the largest possible rank, i.e. E + 2, or ranks are clipped into the last bin);
  * one all-reduce of the dense gradient buffer per step (data-parallel training).
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

    def all_reduce_metrics(self, sums, rr=None):
        """sums: int64 [2, 8] of mre_metrics -> summed over ranks with ONE collective; returns (sums, rr) where rr (float64 [2])
        is rebuilt from the all-reduced fixed-point slot, so every rank -- and every world size -- sees the same bits.
        `rr` (the local float64 sums) is accepted for the old call shape and ignored."""
        sums = sums.clone()
        dist.all_reduce(sums)
        return sums, sums[:, 6].to(torch.float64) / float(1 << 32)

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


class _DevArray:
    """a raw device pointer dressed as a CUDA array (zero-copy torch.as_tensor view of library-owned peer memory)"""

    def __init__(self, ptr, n, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerGroup:
    """mre_peer_group: the flat parameter + gradient buffers of a data-parallel trainer in NVLink peer memory, with the gradient
    reduction fused into the SGD kernel (mre_dp_sgd_step) and the int64 metric all-reduce (mre_peer_allreduce_i64) -- no
    collective library on the data path; torch.distributed only carries the 64-byte IPC handles once, at construction.

        pg = PeerGroup(ctx, n_floats)          # collective over the default process group (or world = 1 without one)
        pg.weights, pg.grads                   # float32 [n_floats] views of this rank's region
        pg.sgd_step(lr / pg.world)             # every rank, once per step

    `peers` (test hook): a list of PeerGroup-s of THIS process standing in for the ranks (connect_local)."""

    def __init__(self, ctx, n_floats, rank=None, world=None, local=False):
        from . import _lib as L
        import ctypes as C
        self.ctx, self.L = ctx, L
        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
            world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank, self.world, self.n = int(rank), int(world), int(n_floats)
        handle = C.create_string_buffer(L.PEER_HANDLE_BYTES)
        out = C.c_void_p()
        L.check(L.lib().mre_peer_group_create(ctx._h, self.rank, self.world, self.n, C.byref(out), handle))
        self._h = out.value
        self.handle = handle.raw
        if not local:
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, self.handle)
            else:
                handles = [self.handle]
            L.check(L.lib().mre_peer_group_connect(self._h, b"".join(handles)))
            self._views()

    @staticmethod
    def connect_local(groups):
        import ctypes as C
        arr = (C.c_void_p * len(groups))(*[g._h for g in groups])
        for g in groups:
            g.L.check(g.L.lib().mre_peer_group_connect_local(g._h, arr))
            g._views()

    def _views(self):
        lib = self.L.lib()
        dev = torch.device("cuda", self.ctx.device)
        self.weights = torch.as_tensor(_DevArray(lib.mre_peer_weights(self._h), self.n), device=dev)
        self.grads = torch.as_tensor(_DevArray(lib.mre_peer_grads(self._h), self.n), device=dev)

    def sgd_step(self, lr, max_blocks=0):
        self.L.check(self.L.lib().mre_dp_sgd_step(self.ctx._h, self._h, float(lr), int(max_blocks), torch.cuda.current_stream().cuda_stream))

    def all_reduce_i64(self, vec):
        assert vec.is_cuda and vec.dtype == torch.int64 and vec.is_contiguous() and vec.numel() <= 64
        self.L.check(self.L.lib().mre_peer_allreduce_i64(self.ctx._h, self._h, vec.data_ptr(), vec.numel(), torch.cuda.current_stream().cuda_stream))
        return vec

    def check(self):
        self.L.check(self.L.lib().mre_peer_group_error(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self.weights = self.grads = None
            self.L.lib().mre_peer_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


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
