"""ctypes/numpy face of oracle/kge_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package (mre_b200) never does.  See kge_oracle.c for the reference
file:line each function restates and for how the restatement is pinned.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build():
    """Compile kge_oracle.c (and, when /root/reference is present, oracle/_ref/Base.so)."""
    subprocess.run(["make", "-s", "-C", _HERE, "all"], check=True)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(_HERE, "kge_oracle.c")
    if not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        build()
    L = C.CDLL(_LIB_PATH)
    L.orc_index_create.restype = C.c_void_p
    L.orc_index_create.argtypes = [C.c_int64, C.c_int64] + [_i64p, _i64p, _i64p, C.c_int64] * 3
    L.orc_index_destroy.argtypes = [C.c_void_p]
    for name in ("orc_train_total", "orc_test_total", "orc_valid_total", "orc_triple_total"):
        getattr(L, name).restype = C.c_int64
        getattr(L, name).argtypes = [C.c_void_p]
    L.orc_get_test.argtypes = [C.c_void_p, _i64p, _i64p, _i64p]
    L.orc_get_train.argtypes = [C.c_void_p, _i64p, _i64p, _i64p]
    L.orc_get_means.argtypes = [C.c_void_p, _f32p, _f32p]
    L.orc_find.restype = C.c_int
    L.orc_find.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64]
    L.orc_rank_from_scores.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int64, C.c_int64, C.c_int64,
                                       C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.orc_rank_from_scores_constrained.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int64, C.c_int64, C.c_int64, _i64p, C.c_int64,
                                                   C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.orc_metrics_add.argtypes = [_f32p, C.c_int64, C.c_int64]
    L.orc_metrics_final.argtypes = [_f32p, _f32p, C.c_int64, _f32p]
    L.orc_corrupt_head.restype = C.c_int64
    L.orc_corrupt_head.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64]
    L.orc_corrupt_tail.restype = C.c_int64
    L.orc_corrupt_tail.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64]
    L.orc_sample_lcg.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                 C.c_int, C.c_int, _i64p, _i64p, _i64p, _f32p]
    L.orc_sample_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int64, C.c_int64,
                                    C.c_int, C.c_int, _i64p, _i64p, _i64p, _f32p]
    _i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    L.orc_sample_subgraph_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, _i64p, _i64p, _i64p, C.c_int64,
                                             _i64p, C.c_int64, _i64p, C.c_int64, C.c_int64, C.c_int, C.c_int, _i32p, _i32p, _i32p]
    L.orc_corrupt_typed_words.restype = C.c_int64
    L.orc_corrupt_typed_words.argtypes = [C.c_void_p, _i64p, _i64p, _i64p, _i64p, C.c_int64, _i64p, C.c_int64, C.POINTER(C.c_uint64), _i64p]
    L.orc_corrupt_typed_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, _i64p, _i64p, _i64p, _i64p, C.c_int64, _i64p]
    L.orc_count_train_leaks.restype = C.c_int64
    L.orc_count_train_leaks.argtypes = [C.c_void_p, _i64p, _i64p, _i64p, C.c_int64, C.c_int64]
    L.orc_philox_selftest.restype = C.c_int
    L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.orc_l2_normalize_rows.argtypes = [_f32p, C.c_int64, C.c_int64, _f32p]
    L.orc_transe_scores.argtypes = [_f32p, _f32p, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                    C.c_int64, C.c_int64, C.c_int64, _f32p]
    L.orc_distmult_scores.argtypes = [_f32p, _f32p, C.c_int64, C.c_int64, C.c_int,
                                      C.c_int64, C.c_int64, C.c_int64, _f32p]
    L.orc_complex_scores.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int64, C.c_int64, C.c_int,
                                     C.c_int64, C.c_int64, C.c_int64, _f32p]
    L.orc_complex_scores_contracted.argtypes = L.orc_complex_scores.argtypes
    L.orc_rank_ties_half.restype = C.c_int64
    L.orc_rank_ties_half.argtypes = [_f32p, C.c_int64]
    L.orc_margin_loss.restype = C.c_double
    L.orc_margin_loss.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_float]
    _lib = L
    return L


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class OracleIndex:
    """Reader.h tables over (train, valid, test) id triples given as (h, t, r) column arrays."""

    def __init__(self, E, R, train, valid, test):
        L = lib()
        self.E, self.R = int(E), int(R)
        cols = []
        for split in (train, valid, test):
            h, t, r = (_i64(x) for x in split)
            cols += [h, t, r, len(h)]
        self._h = C.c_void_p(L.orc_index_create(self.E, self.R, *cols))
        self.train_total = L.orc_train_total(self._h)
        self.test_total = L.orc_test_total(self._h)
        self.valid_total = L.orc_valid_total(self._h)
        self.triple_total = L.orc_triple_total(self._h)

    def __del__(self):
        try:
            lib().orc_index_destroy(self._h)
        except Exception:
            pass

    def _triples(self, fn, n):
        h, t, r = (np.empty(n, np.int64) for _ in range(3))
        fn(self._h, h, t, r)
        return h, t, r

    def test_triples(self):
        """(h, t, r) of the test list in the order the reference iterates it: sorted by (r, h, t)."""
        return self._triples(lib().orc_get_test, self.test_total)

    def train_triples(self):
        return self._triples(lib().orc_get_train, self.train_total)

    def means(self):
        lm, rm = np.empty(self.R, np.float32), np.empty(self.R, np.float32)
        lib().orc_get_means(self._h, lm, rm)
        return lm, rm

    def find(self, h, t, r):
        return bool(lib().orc_find(self._h, int(h), int(t), int(r)))

    def rank_from_scores(self, con, side, h, t, r):
        raw, filt = C.c_int64(), C.c_int64()
        lib().orc_rank_from_scores(self._h, _f32(con), int(side), int(h), int(t), int(r), C.byref(raw), C.byref(filt))
        return raw.value, filt.value

    def rank_from_scores_constrained(self, con, side, h, t, r, type_sorted):
        """Test.h:88-98: counts restricted to the relation's type-constraint candidate list (sorted)"""
        raw, filt = C.c_int64(), C.c_int64()
        ts = _i64(type_sorted)
        lib().orc_rank_from_scores_constrained(self._h, _f32(con), int(side), int(h), int(t), int(r), ts, len(ts),
                                               C.byref(raw), C.byref(filt))
        return raw.value, filt.value

    def corrupt_head(self, h, r, rnd):
        return lib().orc_corrupt_head(self._h, int(h), int(r), C.c_uint64(int(rnd)))

    def corrupt_tail(self, t, r, rnd):
        return lib().orc_corrupt_tail(self._h, int(t), int(r), C.c_uint64(int(rnd)))

    def _alloc(self, B, neg):
        n = B * (1 + neg)
        return np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.float32)

    def sample_lcg(self, states, B, neg, mode=0, bern=1):
        """Reference sampler with the reference LCG; `states` = per-thread next_random (mutated)."""
        threads = len(states)
        bh, bt, br, by = self._alloc(B, neg)
        per = B // threads if B % threads == 0 else B // threads + 1   # Base.cpp:93-100
        for i in range(threads):
            lef, rig = i * per, min((i + 1) * per, B)
            s = C.c_uint64(int(states[i]))
            if lef < rig:
                lib().orc_sample_lcg(self._h, C.byref(s), lef, rig, B, neg, mode, bern, bh, bt, br, by)
            states[i] = s.value
        return bh, bt, br, by

    def sample_philox(self, seed, step, B, neg, mode=0, bern=1, stream=0):
        bh, bt, br, by = self._alloc(B, neg)
        lib().orc_sample_philox(self._h, C.c_uint64(seed), C.c_uint64(step), C.c_uint32(stream), B, neg, mode, bern,
                                bh, bt, br, by)
        return bh, bt, br, by

    def sample_subgraph_philox(self, seed, step, eh, et, er, nodes, l2g, neg, bern=0, filt=1, stream=0):
        """CPU replay of mre_sample_subgraph -> (edge_index int32 [2, n(1+neg)], edge_type int32 [n(1+neg)])"""
        n = len(eh) * (1 + neg)
        oh, ot, orl = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
        lib().orc_sample_subgraph_philox(self._h, C.c_uint64(seed), C.c_uint64(step), C.c_uint32(stream), _i64(eh), _i64(et),
                                         _i64(er), len(eh), _i64(nodes), len(nodes), _i64(l2g), len(l2g), neg, int(bern),
                                         int(filt), oh, ot, orl)
        return np.stack([oh, ot]), orl

    def corrupt_typed_words(self, tail_ptr, tail_idx, h, r, words, lcg_state=0):
        """corrupt(h, r) of Corrupt.h:179-195 fed the given random words in order (the reference's libc rand() values) and, for
        the corrupt_head fallback, thread 0's LCG started at lcg_state -> (tails, words consumed, final LCG state)"""
        out = np.zeros(len(h), np.int64)
        st = C.c_uint64(int(lcg_state))
        used = lib().orc_corrupt_typed_words(self._h, _i64(tail_ptr), _i64(tail_idx), _i64(h), _i64(r), len(h), _i64(words),
                                             len(words), C.byref(st), out)
        return out, int(used), st.value

    def corrupt_typed_philox(self, seed, step, tail_ptr, tail_idx, h, r, stream=0):
        """CPU replay of mre_corrupt_typed (corrupt(h, r), Corrupt.h:179-195) over the sorted per-relation tail-type lists"""
        out = np.zeros(len(h), np.int64)
        idx = _i64(tail_idx) if len(tail_idx) else np.zeros(1, np.int64)
        lib().orc_corrupt_typed_philox(self._h, C.c_uint64(seed), C.c_uint64(step), C.c_uint32(stream), _i64(tail_ptr), idx,
                                       _i64(h), _i64(r), len(h), out)
        return out

    def count_train_leaks(self, bh, bt, br, start, stop):
        return lib().orc_count_train_leaks(self._h, _i64(bh), _i64(bt), _i64(br), start, stop)


class MetricAccumulator:
    """Test.h's float32 accumulators + finaliser; result order (mrr, mr, hit10, hit3, hit1)."""

    def __init__(self):
        self.head = np.zeros(10, np.float32)
        self.tail = np.zeros(10, np.float32)

    def add(self, side, raw, filt):
        lib().orc_metrics_add(self.head if side == 0 else self.tail, int(raw), int(filt))

    def final(self, test_total):
        out = np.zeros(5, np.float32)
        lib().orc_metrics_final(self.head, self.tail, int(test_total), out)
        return tuple(float(x) for x in out)


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return tuple(o)


def l2_normalize_rows(x):
    x = _f32(x)
    out = np.empty_like(x)
    lib().orc_l2_normalize_rows(x, x.shape[0], x.shape[1], out)
    return out


def transe_scores(ent, rel, p_norm, side, h, t, r):
    ent, rel = _f32(ent), _f32(rel)
    out = np.empty(ent.shape[0], np.float32)
    lib().orc_transe_scores(ent, rel, ent.shape[0], ent.shape[1], p_norm, side, h, t, r, out)
    return out


def distmult_scores(ent, rel, side, h, t, r):
    ent, rel = _f32(ent), _f32(rel)
    out = np.empty(ent.shape[0], np.float32)
    lib().orc_distmult_scores(ent, rel, ent.shape[0], ent.shape[1], side, h, t, r, out)
    return out


def complex_scores(ent_re, ent_im, rel_re, rel_im, side, h, t, r):
    ent_re, ent_im, rel_re, rel_im = _f32(ent_re), _f32(ent_im), _f32(rel_re), _f32(rel_im)
    out = np.empty(ent_re.shape[0], np.float32)
    lib().orc_complex_scores(ent_re, ent_im, rel_re, rel_im, ent_re.shape[0], ent_re.shape[1], side, h, t, r, out)
    return out


def complex_scores_contracted(ent_re, ent_im, rel_re, rel_im, side, h, t, r):
    """ComplEx in contraction form ([a;b] . [e_re;e_im], K = 2D, sequential float32): the tcgen05 path's association"""
    ent_re, ent_im, rel_re, rel_im = _f32(ent_re), _f32(ent_im), _f32(rel_re), _f32(rel_im)
    out = np.empty(ent_re.shape[0], np.float32)
    lib().orc_complex_scores_contracted(ent_re, ent_im, rel_re, rel_im, ent_re.shape[0], ent_re.shape[1], side, h, t, r, out)
    return out


def rank_ties_half(scores):
    s = _f32(scores)
    return lib().orc_rank_ties_half(s, len(s))


def margin_loss(score, B, neg, margin):
    return lib().orc_margin_loss(_f32(score), B, neg, margin)
