"""Host-side objects over the C ABI: the knowledge-graph index, the per-device workspace and the rank /
sample / train-step calls on torch CUDA tensors.  torch supplies device memory and streams only; every
computation below happens inside libmre_b200.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _ptr(t):
    """device/host pointer of a contiguous torch tensor or numpy array (None -> NULL)"""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        assert t.flags["C_CONTIGUOUS"]
        return t.ctypes.data
    assert t.is_contiguous()
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class KGIndex:
    """Reader.h's tables (OpenKE/openke/base/Reader.h:53-257) built by mre_index_create*."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)
        self.device = None
        lib = L.lib()
        self.ent_tot = lib.mre_index_total(self._h, L.TOTAL_ENTITY)
        self.rel_tot = lib.mre_index_total(self._h, L.TOTAL_RELATION)
        self.train_tot = lib.mre_index_total(self._h, L.TOTAL_TRAIN)
        self.valid_tot = lib.mre_index_total(self._h, L.TOTAL_VALID)
        self.test_tot = lib.mre_index_total(self._h, L.TOTAL_TEST)
        self.triple_tot = lib.mre_index_total(self._h, L.TOTAL_TRIPLE)

    @classmethod
    def from_arrays(cls, E, R, train, valid=None, test=None):
        """train / valid / test: (h, t, r) column arrays (OpenKE's file column order)."""
        empty = (np.zeros(0, np.int64),) * 3
        cols = []
        keep = []
        for split in (train, valid or empty, test or empty):
            h, t, r = (_i64(x) for x in split)
            assert len(h) == len(t) == len(r)
            keep += [h, t, r]
            cols += [h.ctypes.data, t.ctypes.data, r.ctypes.data, len(h)]
        out = C.c_void_p()
        L.check(L.lib().mre_index_create(int(E), int(R), *cols, C.byref(out)))
        return cls(out.value)

    @classmethod
    def from_arrays_device(cls, E, R, train, valid=None, test=None, device=0):
        """from_arrays + to_device with the sorts, the de-duplication and the relation counters run on the GPU
        (mre_index_create_device: LSD radix sort, csrc/index_build.cu); `build_ms` holds the device time of the build."""
        empty = (np.zeros(0, np.int64),) * 3
        cols, keep = [], []
        for split in (train, valid or empty, test or empty):
            h, t, r = (_i64(x) for x in split)
            assert len(h) == len(t) == len(r)
            keep += [h, t, r]
            cols += [h.ctypes.data, t.ctypes.data, r.ctypes.data, len(h)]
        out, ms = C.c_void_p(), C.c_double(0.0)
        L.check(L.lib().mre_index_create_device(int(device), int(E), int(R), *cols, C.byref(out), C.byref(ms)))
        ix = cls(out.value)
        ix.device, ix.build_ms = int(device), ms.value
        return ix

    def device_column(self, which):
        """one column of the device tables as a host array (mre_index_device_column)"""
        lib = L.lib()
        n = lib.mre_index_device_column(self._h, int(which), None)
        if n < 0:
            raise L.MreError(lib.mre_last_error().decode())
        out = np.empty(n, np.float32 if which == 10 else np.int64)
        if lib.mre_index_device_column(self._h, int(which), out.ctypes.data) < 0:
            raise L.MreError(lib.mre_last_error().decode())
        return out

    @classmethod
    def from_dir(cls, in_path, device=None):
        """setInPath + importTrainFiles + importTestFiles (Setting.h:17-27, Reader.h:53-257).  `device`: build the tables on
        that GPU (mre_index_create_from_dir_device: radix sorts instead of the host sorts; the index then lives on it)."""
        out = C.c_void_p()
        if device is None:
            L.check(L.lib().mre_index_create_from_dir(str(in_path).encode(), C.byref(out)))
            return cls(out.value)
        ms = C.c_double(0.0)
        L.check(L.lib().mre_index_create_from_dir_device(str(in_path).encode(), int(device), C.byref(out), C.byref(ms)))
        ix = cls(out.value)
        ix.device, ix.build_ms = int(device), ms.value
        return ix

    def __del__(self):
        try:
            if self._h:
                L.lib().mre_index_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def to_device(self, device=0):
        L.check(L.lib().mre_index_to_device(self._h, int(device)))
        self.device = int(device)
        return self

    def _split(self, which, n):
        h, t, r = (np.empty(n, np.int64) for _ in range(3))
        L.check(L.lib().mre_index_get_split(self._h, which, h.ctypes.data, t.ctypes.data, r.ctypes.data))
        return h, t, r

    def train_triples(self):
        """de-duplicated train list, (h,r,t) order == trainList"""
        return self._split(L.SPLIT_TRAIN, self.train_tot)

    def valid_triples(self):
        return self._split(L.SPLIT_VALID, self.valid_tot)

    def test_triples(self):
        """test list in the order Tester iterates it: sorted (r,h,t) (Reader.h:227)"""
        return self._split(L.SPLIT_TEST, self.test_tot)

    def means(self):
        tph, hpt = np.empty(self.rel_tot, np.float32), np.empty(self.rel_tot, np.float32)
        L.check(L.lib().mre_index_get_means(self._h, tph.ctypes.data, hpt.ctypes.data))
        return tph, hpt

    def find(self, h, t, r):
        return bool(L.lib().mre_index_find(self._h, int(h), int(t), int(r)))

    # ---- type constraints (importTypeFiles, Reader.h:267-317)
    @property
    def has_type_constrain(self):
        return L.lib().mre_index_type_total(self._h, 0) >= 0

    def load_type_constrain(self, path):
        L.check(L.lib().mre_index_load_type_constrain(self._h, str(path).encode()))

    def set_type_constrain(self, head_ptr, head_idx, tail_ptr, tail_idx):
        a = [_i64(x) for x in (head_ptr, head_idx, tail_ptr, tail_idx)]
        L.check(L.lib().mre_index_set_type_constrain(self._h, *(x.ctypes.data for x in a)))

    def type_constrain(self, side):
        """(ptr [R+1], idx) of the head (side 0) / tail (side 1) lists, each relation's slice sorted and unique"""
        n = L.lib().mre_index_type_total(self._h, int(side))
        if n < 0:
            raise L.MreError("the index holds no type constraints (type_constrain.txt missing)")
        ptr, idx = np.empty(self.rel_tot + 1, np.int64), np.empty(max(n, 1), np.int64)
        L.check(L.lib().mre_index_get_type_constrain(self._h, int(side), ptr.ctypes.data, idx.ctypes.data))
        return ptr, idx[:n]


class Context:
    """mre_ctx: scratch buffers and launch accounting for one device."""

    def __init__(self, device=0):
        out = C.c_void_p()
        L.check(L.lib().mre_ctx_create(int(device), C.byref(out)))
        self._h = out
        self.device = int(device)

    def __del__(self):
        try:
            if self._h:
                L.lib().mre_ctx_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def sm_count(self):
        return L.lib().mre_ctx_sm_count(self._h)

    @property
    def launches(self):
        return L.lib().mre_ctx_launch_count(self._h)

    def option(self, key, value):
        """mre_ctx_option: 'bil_products' (1 | 3), 'bil_pair', 'transe_ctas_per_sm', 'zsl_fp32'"""
        L.check(L.lib().mre_ctx_option(self._h, key.encode(), int(value)))
        return self

    def stat(self, key):
        """mre_ctx_stat: read-and-reset a device-side counter ('bil_rescored')"""
        v = C.c_int64()
        L.check(L.lib().mre_ctx_stat(self._h, key.encode(), C.byref(v)))
        return v.value

    def timing(self, enable):
        L.check(L.lib().mre_ctx_timing(self._h, 1 if enable else 0))

    def timing_read(self):
        ms, n = C.c_double(), C.c_int64()
        L.check(L.lib().mre_ctx_timing_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def probe_fp32_peak(self):
        v = C.c_double()
        L.check(L.lib().mre_probe_fp32_peak(self._h, C.byref(v)))
        return v.value

    def probe_mufu_peak(self):
        v = C.c_double()
        L.check(L.lib().mre_probe_mufu_peak(self._h, C.byref(v)))
        return v.value

    def probe_bf16_peak(self):
        v = C.c_double()
        L.check(L.lib().mre_probe_bf16_peak(self._h, C.byref(v)))
        return v.value

    def probe_tf32_peak(self):
        v = C.c_double()
        L.check(L.lib().mre_probe_tf32_peak(self._h, C.byref(v)))
        return v.value


SCORERS = {"transe": L.TRANSE, "distmult": L.DISTMULT, "complex": L.COMPLEX, "rotate": L.ROTATE}
RANK_MODES = {"strict": L.RANK_STRICT, "ties_half": L.RANK_TIES_HALF, "pessimistic": L.RANK_PESSIMISTIC}


class CandidateGroups:
    """Candidate lists shared by runs of queries (rel2candidates, utils/gen_rel2candidates.py:23-27).
    qptr / cptr: host int64 prefix arrays [G+1]; cand: device int64, each slice sorted and unique."""

    def __init__(self, qptr, cptr, cand):
        self.qptr = _i64(qptr)
        self.cptr = _i64(cptr)
        self.cand = cand
        assert len(self.qptr) == len(self.cptr)

    @classmethod
    def from_lists(cls, query_counts, cand_lists, device):
        qptr = np.concatenate([[0], np.cumsum(query_counts)])
        uniq = [np.unique(_i64(c)) for c in cand_lists]
        cptr = np.concatenate([[0], np.cumsum([len(u) for u in uniq])])
        flat = np.concatenate(uniq) if uniq else np.zeros(0, np.int64)
        return cls(qptr, cptr, torch.from_numpy(flat).to(device))


class Ranker:
    """Fused score + filtered-rank engine for one device (mre_rank / mre_rank_host / mre_metrics)."""

    def __init__(self, ctx=None, device=0):
        self.ctx = ctx or Context(device)
        self.device = torch.device("cuda", self.ctx.device)

    def _job(self, scorer, tables, Q, side, p_norm, normalize, index, filt, groups, filt_csr, phase_div=0.0):
        job = L.RankJob()
        if scorer == "complex":
            ent, ent_im, rel, rel_im = tables
            job.ent_im, job.rel_im = _ptr(ent_im), _ptr(rel_im)
            assert ent_im.shape == ent.shape and rel_im.shape == rel.shape
        else:
            ent, rel = tables
        for t in tables:
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), "tables must be contiguous float32 CUDA tensors"
        job.ent, job.rel = _ptr(ent), _ptr(rel)
        job.E, job.D = ent.shape
        job.R = rel.shape[0]
        if scorer == "rotate":                  # ent rows [re | im], rel rows = phases (RotatE.py:13-14)
            assert 2 * rel.shape[1] == ent.shape[1] and phase_div > 0
            job.rotate_phase_div = float(phase_div)
        else:
            assert rel.shape[1] == ent.shape[1]
        job.scorer = SCORERS[scorer]
        job.p_norm, job.normalize = int(p_norm), int(bool(normalize))
        if filt is None:
            filt = "csr" if filt_csr is not None else ("index" if index is not None else "none")
        job.filter = {"none": L.FILTER_NONE, "index": L.FILTER_INDEX, "csr": L.FILTER_CSR}[filt]
        job.Q = int(Q)
        keep = []
        if isinstance(side, (int, np.integer)):
            job.side = int(side)
        else:
            job.side = 0
            keep.append(side)
            job.q_side = _ptr(side)
        if groups is not None:
            job.n_groups = len(groups.qptr) - 1
            job.group_qptr, job.group_cptr = groups.qptr.ctypes.data, groups.cptr.ctypes.data
            job.cand_idx = _ptr(groups.cand)
        if filt_csr is not None:
            job.filt_ptr, job.filt_idx = _ptr(filt_csr[0]), _ptr(filt_csr[1])
            # (ptr, idx) or (ptr, idx, nnz): nnz = filt_ptr[Q] (an upper bound will do; 0 = unknown to the caller)
            job.filt_nnz = int(filt_csr[2]) if len(filt_csr) > 2 else int(filt_csr[1].numel())
        return job, keep

    def rank(self, scorer, tables, q_h, q_t, q_r, side, *, p_norm=1, normalize=False, index=None, filter=None,
             groups=None, filt_csr=None, out=None, phase_div=0.0):
        """Device queries -> device int32 counts [4, Q] = (raw_lt, raw_eq, filt_lt, filt_eq).  Asynchronous.
        `phase_div`: RotatE only (rel_embedding_range / pi)."""
        Q = q_h.numel()
        job, keep = self._job(scorer, tables, Q, side, p_norm, normalize, index, filter, groups, filt_csr, phase_div)
        for t in (q_h, q_t, q_r):
            assert t.is_cuda and t.dtype == torch.int64 and t.is_contiguous() and t.numel() == Q
        if not isinstance(side, (int, np.integer)):
            assert side.is_cuda and side.dtype == torch.uint8 and side.numel() == Q
        job.q_h, job.q_t, job.q_r = _ptr(q_h), _ptr(q_t), _ptr(q_r)
        counts = out if out is not None else torch.empty((4, Q), dtype=torch.int32, device=self.device)
        job.counts = _ptr(counts)
        L.check(L.lib().mre_rank(self.ctx._h, index._h if index is not None else None, C.byref(job), _stream()))
        return counts

    def rank_host(self, scorer, tables, q_h, q_t, q_r, side, *, p_norm=1, normalize=False, index=None, filter=None,
                  groups=None, filt_csr=None, out=None, phase_div=0.0):
        """Host (numpy / pinned torch CPU) queries -> host int32 counts [4, Q]; copies included; synchronous."""
        Q = len(q_h)
        job, keep = self._job(scorer, tables, Q, side, p_norm, normalize, index, filter, groups, filt_csr, phase_div)
        arrs = []
        for a in (q_h, q_t, q_r):
            a = a if isinstance(a, torch.Tensor) else _i64(a)
            assert (a.dtype == torch.int64) if isinstance(a, torch.Tensor) else True
            arrs.append(a)
        job.q_h, job.q_t, job.q_r = (_ptr(a) for a in arrs)
        if not isinstance(side, (int, np.integer)):
            side = side if isinstance(side, torch.Tensor) else np.ascontiguousarray(side, dtype=np.uint8)
            job.q_side = _ptr(side)
        counts = out if out is not None else np.empty((4, Q), np.int32)
        job.counts = _ptr(counts)
        L.check(L.lib().mre_rank_host(self.ctx._h, index._h if index is not None else None, C.byref(job), _stream()))
        return counts

    def predict(self, scorer, tables, q_h, q_t, q_r, side, query=0, *, p_norm=1, normalize=False):
        """Model.predict's float32[E] vector for one query (device tensor)."""
        if scorer == "rotate":
            raise L.MreError("RotatE is served by rank() only: use the model's own predict() for explicit batches")
        Q = q_h.numel()
        job, keep = self._job(scorer, tables, Q, side, p_norm, normalize, None, "none", None, None)
        job.q_h, job.q_t, job.q_r = _ptr(q_h), _ptr(q_t), _ptr(q_r)
        out = torch.empty(job.E, dtype=torch.float32, device=self.device)
        L.check(L.lib().mre_predict(self.ctx._h, C.byref(job), int(query), _ptr(out), _stream()))
        return out

    def bilinear_scores(self, scorer, tables, q_h, q_t, q_r, side):
        """[Q, E] raw tensor-core similarities of a DistMult / ComplEx all-entity job (mre_bilinear_scores)"""
        Q = q_h.numel()
        job, keep = self._job(scorer, tables, Q, side, 1, False, None, "none", None, None)
        job.q_h, job.q_t, job.q_r = _ptr(q_h), _ptr(q_t), _ptr(q_r)
        out = torch.empty((Q, job.E), dtype=torch.float32, device=self.device)
        L.check(L.lib().mre_bilinear_scores(self.ctx._h, C.byref(job), _ptr(out), _stream()))
        return out

    def metrics(self, counts, side, rank_mode="strict", raw=False, hist_len=0):
        """counts [4, Q] (device) -> dict of per-side integer sums + float64 reciprocal-rank sums (+ histogram)."""
        Q = counts.shape[1]
        sums = torch.empty((2, 8), dtype=torch.int64, device=self.device)     # mre_metrics writes all 16 + 2 slots
        rr = torch.empty(2, dtype=torch.float64, device=self.device)
        hist = torch.zeros(hist_len, dtype=torch.int64, device=self.device) if hist_len else None
        if isinstance(side, (int, np.integer)):
            side_ptr, side_val = None, int(side)
        else:
            side_ptr, side_val = _ptr(side), 0
        L.check(L.lib().mre_metrics(self.ctx._h, _ptr(counts), side_ptr, side_val, Q, RANK_MODES[rank_mode], int(raw),
                                    _ptr(sums), _ptr(rr), _ptr(hist), int(hist_len), _stream()))
        return {"sums": sums, "rr": rr, "hist": hist}


RR_FIXED_ONE = float(1 << 32)     # sums[s][6] is the reciprocal-rank sum in 32.32 fixed point (metrics.cu)


def summarize(sums, rr=None):
    """per-side integer sums (host numpy [2, 8]) -> dict of means per side: mrr, mr, hits@1/3/5/10.  MRR comes from the
    INTEGER fixed-point reciprocal-rank sum (slot 6), so the tuple is bit-identical however the queries were sharded; pass
    the float64 `rr` sums to use those instead (one device, <= 2^-32 apart)."""
    out = []
    for s in range(2):
        n = int(sums[s][0])
        if n == 0:
            out.append(None)
            continue
        rr_s = float(rr[s]) if rr is not None else float(int(sums[s][6])) / RR_FIXED_ONE
        out.append({"n": n, "mr": float(sums[s][1]) / n, "mrr": rr_s / n, "hits1": float(sums[s][2]) / n,
                    "hits3": float(sums[s][3]) / n, "hits5": float(sums[s][4]) / n, "hits10": float(sums[s][5]) / n})
    return out


class Sampler:
    """mre_sample: Philox Bernoulli negative sampler (Base.cpp:78-197)."""

    def __init__(self, index, ctx=None, seed=0, stream_id=0):
        self.index = index
        self.ctx = ctx or Context(index.device if index.device is not None else 0)
        if index.device is None:
            index.to_device(self.ctx.device)
        self.seed, self.stream_id = int(seed), int(stream_id)
        self.device = torch.device("cuda", self.ctx.device)

    def sample(self, step, B, neg, mode=0, bern=1, out=None):
        """device batch (h, t, r, y) of B (1 + neg) rows; `out` = four preallocated device tensors to write into (a caller that
        samples on a side stream keeps its own buffers so that the caching allocator never crosses streams)"""
        n = B * (1 + neg)
        if out is not None:
            h, t, r, y = out
            assert all(x.is_cuda and x.numel() == n for x in out) and h.dtype == torch.int64 and y.dtype == torch.float32
        else:
            h, t, r = (torch.empty(n, dtype=torch.int64, device=self.device) for _ in range(3))
            y = torch.empty(n, dtype=torch.float32, device=self.device)
        L.check(L.lib().mre_sample(self.ctx._h, self.index._h, self.seed, int(step), self.stream_id, B, neg, mode, bern,
                                   _ptr(h), _ptr(t), _ptr(r), _ptr(y), _stream()))
        return h, t, r, y

    def sample_host(self, step, B, neg, mode=0, bern=1, out=None):
        n = B * (1 + neg)
        if out is None:
            out = (np.empty(n, np.int64), np.empty(n, np.int64), np.empty(n, np.int64), np.empty(n, np.float32))
        h, t, r, y = out
        L.check(L.lib().mre_sample_host(self.ctx._h, self.index._h, self.seed, int(step), self.stream_id, B, neg, mode, bern,
                                        _ptr(h), _ptr(t), _ptr(r), _ptr(y), _stream()))
        return h, t, r, y

    def corrupt_typed(self, step, h, r):
        """mre_corrupt_typed: corrupt(h, r) of Corrupt.h:179-195 for device int64 arrays h, r -> device int64 tails"""
        out = torch.empty_like(h)
        L.check(L.lib().mre_corrupt_typed(self.ctx._h, self.index._h, self.seed, int(step), self.stream_id, _ptr(h), _ptr(r),
                                          h.numel(), _ptr(out), _stream()))
        return out


def transe_margin_step(ctx, ent, rel, h, t, r, B, neg, margin, p_norm=1, normalize=True, grad_ent=None, grad_rel=None,
                       want_scores=False):
    """mre_transe_margin_step on device tensors -> (loss[1], grad_ent, grad_rel, scores|None)."""
    E, D = ent.shape
    R = rel.shape[0]
    if grad_ent is None:
        grad_ent = torch.zeros_like(ent)
    if grad_rel is None:
        grad_rel = torch.zeros_like(rel)
    loss = torch.zeros(1, dtype=torch.float32, device=ent.device)
    scores = torch.empty(B * (1 + neg), dtype=torch.float32, device=ent.device) if want_scores else None
    L.check(L.lib().mre_transe_margin_step(ctx._h, _ptr(ent), _ptr(rel), E, R, D, _ptr(h), _ptr(t), _ptr(r), B, neg,
                                           float(margin), int(p_norm), int(bool(normalize)), _ptr(grad_ent), _ptr(grad_rel),
                                           _ptr(loss), _ptr(scores), _stream()))
    return loss, grad_ent, grad_rel, scores


def ns_train_step(ctx, scorer, tables, grads, h, t, r, B, neg, loss_kind, margin=0.0, adv=0, temperature=0.0, p_norm=1,
                  normalize=False, want_scores=False):
    """mre_ns_train_step: fused score + negative-sampling loss + backward for any scorer; `grads` are accumulated into.
    tables / grads: (ent, rel) or, for ComplEx, (ent_re, ent_im, rel_re, rel_im).  -> loss[1] (and the n scores if asked)."""
    if scorer == "complex":
        ent, ent_im, rel, rel_im = tables
        g_ent, g_ent_im, g_rel, g_rel_im = grads
    else:
        (ent, rel), ent_im, rel_im = tables, None, None
        (g_ent, g_rel), g_ent_im, g_rel_im = grads, None, None
    loss = torch.empty(1, dtype=torch.float32, device=ent.device)
    scores = torch.empty(B * (1 + neg), dtype=torch.float32, device=ent.device) if want_scores else None
    L.check(L.lib().mre_ns_train_step(ctx._h, SCORERS[scorer], _ptr(ent), _ptr(ent_im), _ptr(rel), _ptr(rel_im), ent.shape[1],
                                      _ptr(h), _ptr(t), _ptr(r), B, neg, int(loss_kind), float(margin), int(adv), float(temperature),
                                      int(p_norm), int(bool(normalize)), _ptr(g_ent), _ptr(g_ent_im), _ptr(g_rel), _ptr(g_rel_im),
                                      _ptr(loss), _ptr(scores), _stream()))
    return (loss, scores) if want_scores else loss


def sgd_update(ctx, w, g, lr):
    L.check(L.lib().mre_sgd_update(ctx._h, _ptr(w), _ptr(g), w.numel(), float(lr), _stream()))
