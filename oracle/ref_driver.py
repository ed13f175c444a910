"""ctypes driver for the UNMODIFIED reference library oracle/_ref/Base.so -- TEST INFRASTRUCTURE ONLY.

oracle/_ref/Base.so is compiled by oracle/Makefile from /root/reference/OpenKE/openke/base/Base.cpp
(where the sources lie; nothing is copied).  This module re-creates, outside the product, the ctypes
contract of the loaders the reference lost to its .gitignore (openke/data/{Train,Test}DataLoader.py;
SURVEY Appendix B), derived from the C signatures:
  sampling(...)                  OpenKE/openke/base/Base.cpp:161-174
  getHeadBatch/getTailBatch      OpenKE/openke/base/Test.h:36-53
  testHead/testTail              OpenKE/openke/base/Test.h:65-192
  test_link_prediction + getters OpenKE/openke/base/Test.h:232-390  (argtypes as Tester.py:22-36)
Base.so keeps all state in globals: ONE dataset per process.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "Base.so")


def available():
    return os.path.exists(REF_LIB)


def write_type_constrain(path, R, head_lists, tail_lists):
    """type_constrain.txt in the format importTypeFiles reads (Reader.h:267-317; produced by benchmarks/*/n-n.py):
    a count line, then per relation "rel n ids..." for the heads and the same for the tails."""
    with open(os.path.join(path, "type_constrain.txt"), "w") as f:
        f.write(f"{R}\n")
        for r in range(R):
            for lst in (head_lists[r], tail_lists[r]):
                f.write("\t".join([str(r), str(len(lst))] + [str(int(x)) for x in lst]) + "\n")


def read_type_constrain(path):
    """-> (head_lists, tail_lists): dict rel -> sorted unique int64 array"""
    toks = open(os.path.join(path, "type_constrain.txt")).read().split()
    n, pos = int(toks[0]), 1
    heads, tails = {}, {}
    for _ in range(n):
        for dst in (heads, tails):
            rel, tot = int(toks[pos]), int(toks[pos + 1])
            dst[rel] = np.unique(np.asarray(toks[pos + 2:pos + 2 + tot], dtype=np.int64))
            pos += 2 + tot
    return heads, tails


def write_benchmark_dir(path, E, R, train, valid, test):
    """Materialise OpenKE's text format (count line, then 'h t r' rows: OpenKE/README.md:126-141)."""
    os.makedirs(path, exist_ok=True)
    for name, n in (("entity2id.txt", E), ("relation2id.txt", R)):
        with open(os.path.join(path, name), "w") as f:
            f.write(f"{n}\n")
            f.write("".join(f"x{i}\t{i}\n" for i in range(n)))
    for name, (h, t, r) in (("train2id.txt", train), ("valid2id.txt", valid), ("test2id.txt", test)):
        arr = np.stack([np.asarray(h), np.asarray(t), np.asarray(r)], 1).astype(np.int64)
        with open(os.path.join(path, name), "w") as f:
            f.write(f"{len(arr)}\n")
            np.savetxt(f, arr, fmt="%d")
    return path


class RefOpenKE:
    def __init__(self, in_path, threads=1, bern=1, libc_seed=1):
        if not in_path.endswith("/"):
            in_path += "/"
        L = C.CDLL(REF_LIB)
        self.L = L
        L.sampling.argtypes = [C.c_void_p] * 4 + [C.c_int64] * 7
        L.getHeadBatch.argtypes = [C.c_void_p] * 3
        L.getTailBatch.argtypes = [C.c_void_p] * 3
        L.testHead.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.testTail.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.test_link_prediction.argtypes = [C.c_int64]
        for g in ("getTestLinkMRR", "getTestLinkMR", "getTestLinkHit10", "getTestLinkHit3", "getTestLinkHit1"):
            getattr(L, g).argtypes = [C.c_int64]
            getattr(L, g).restype = C.c_float
        for g in ("getEntityTotal", "getRelationTotal", "getTripleTotal", "getTrainTotal", "getTestTotal",
                  "getValidTotal", "getWorkThreads"):
            getattr(L, g).restype = C.c_int64
        L.setWorkThreads.argtypes = [C.c_int64]
        L.setBern.argtypes = [C.c_int64]
        self._path = C.create_string_buffer(in_path.encode(), len(in_path) * 2)
        L.setInPath(self._path)
        L.setBern(bern)
        L.setWorkThreads(threads)
        # randReset seeds each thread's LCG from libc rand() (Random.h:11-15); pin libc's state first
        C.CDLL(None).srand(libc_seed)
        L.randReset()
        L.importTrainFiles()
        self.threads = threads
        self.ent_tot = L.getEntityTotal()
        self.rel_tot = L.getRelationTotal()
        self.train_tot = L.getTrainTotal()
        self._test_loaded = False

    @staticmethod
    def lcg_seeds(threads, libc_seed=1):
        """The seeds randReset() hands out: successive libc rand() values after srand(libc_seed)."""
        libc = C.CDLL(None)
        libc.srand(libc_seed)
        out = [libc.rand() for _ in range(threads)]
        return out

    def sampling(self, B, neg, mode=0):
        n = B * (1 + neg)
        h, t, r = (np.zeros(n, np.int64) for _ in range(3))
        y = np.zeros(n, np.float32)
        self.L.sampling(h.ctypes.data, t.ctypes.data, r.ctypes.data, y.ctypes.data, B, neg, 0, mode, 1, 0, 0)
        return h, t, r, y

    def load_test(self, type_files=False):
        if not self._test_loaded:
            self.L.importTestFiles()
            if type_files:
                self.L.importTypeFiles()      # Reader.h:267-317 (needs type_constrain.txt in the benchmark directory)
            self._test_loaded = True
        self.test_tot = self.L.getTestTotal()
        self.L.initTest()
        E = self.ent_tot
        self._bh, self._bt, self._br = (np.zeros(E, np.int64) for _ in range(3))

    def head_batch(self):
        self.L.getHeadBatch(self._bh.ctypes.data, self._bt.ctypes.data, self._br.ctypes.data)
        return {"batch_h": self._bh, "batch_t": self._bt[:1], "batch_r": self._br[:1], "mode": "head_batch"}

    def tail_batch(self):
        self.L.getTailBatch(self._bh.ctypes.data, self._bt.ctypes.data, self._br.ctypes.data)
        return {"batch_h": self._bh[:1], "batch_t": self._bt, "batch_r": self._br[:1], "mode": "tail_batch"}

    def test_head(self, score, index, type_constrain=0):
        score = np.ascontiguousarray(score, np.float32)
        self.L.testHead(score.ctypes.data, index, type_constrain)

    def test_tail(self, score, index, type_constrain=0):
        score = np.ascontiguousarray(score, np.float32)
        self.L.testTail(score.ctypes.data, index, type_constrain)

    def finish(self, type_constrain=0):
        """-> (mrr, mr, hit10, hit3, hit1), Tester.py:83-91.  (The C side printf's its table.)"""
        L = self.L
        L.test_link_prediction(type_constrain)
        tc = type_constrain
        return (L.getTestLinkMRR(tc), L.getTestLinkMR(tc), L.getTestLinkHit10(tc), L.getTestLinkHit3(tc),
                L.getTestLinkHit1(tc))
