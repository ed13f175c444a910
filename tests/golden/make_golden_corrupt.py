"""golden_corrupt.npz -- the compiled reference's own corrupt(h, r) (OpenKE/openke/base/Corrupt.h:179-195) on FB15K237.

corrupt() is not extern "C" but Base.so exports it (_Z7corruptll).  It draws with libc rand() (Random.h:32-34), so after
srand(k) its outputs are a function of the rand() stream (and, for the corrupt_head fallback after 1000 failed draws, of thread 0's
LCG, whose seed randReset took from rand() after srand(1)): the script records the pairs, the reference's tails, the rand()
values it consumed and the LCG seed, and asserts that oracle/kge_oracle.c's restatement fed the same values reproduces every tail.
Run here (needs /root/reference and oracle/_ref/Base.so):  python tests/golden/make_golden_corrupt.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import golden_util as gu  # noqa: E402
from oracle import kge_oracle as ko, ref_driver as rd  # noqa: E402

BENCH = "/root/reference/OpenKE/benchmarks/FB15K237/"
LIBC_SEED, N_TRAIN_PAIRS, N_RANDOM_PAIRS = 7, 1536, 512


def raw_tail_lists(path, R):
    """the tail lists exactly as importTypeFiles keeps them: sorted, repeated ids kept (Reader.h:303-312)"""
    toks = open(os.path.join(path, "type_constrain.txt")).read().split()
    n, pos, tails = int(toks[0]), 1, {}
    for _ in range(n):
        for keep in (False, True):
            rel, tot = int(toks[pos]), int(toks[pos + 1])
            if keep:
                tails[rel] = np.sort(np.asarray(toks[pos + 2:pos + 2 + tot], dtype=np.int64))
            pos += 2 + tot
    return [tails.get(r, np.zeros(0, np.int64)) for r in range(R)]


def main():
    ids = gu.load("fb15k237_ids.npz")
    E, R = int(ids["E"]), int(ids["R"])
    tr, va, te = (gu.split_cols(ids, s) for s in ("train", "valid", "test"))
    ref = rd.RefOpenKE(BENCH, threads=1)
    ref.load_test(type_files=True)
    corrupt = ref.L._Z7corruptll
    corrupt.argtypes, corrupt.restype = [C.c_int64, C.c_int64], C.c_int64
    rng = np.random.default_rng(11)
    pick = rng.choice(len(tr[0]), N_TRAIN_PAIRS, replace=False)
    h = np.concatenate([tr[0][pick], rng.integers(0, E, N_RANDOM_PAIRS)]).astype(np.int64)
    r = np.concatenate([tr[2][pick], rng.integers(0, R, N_RANDOM_PAIRS)]).astype(np.int64)
    tails = raw_tail_lists(BENCH, R)
    keep = np.array([len(tails[x]) > 0 for x in r])            # rand() % 0 in the reference: not a defined input
    h, r = h[keep], r[keep]
    libc = C.CDLL(None)
    libc.srand(LIBC_SEED)
    want = np.array([corrupt(int(a), int(b)) for a, b in zip(h, r)], np.int64)
    libc.srand(LIBC_SEED)
    words = np.array([libc.rand() for _ in range(256 * len(h))], np.int64)
    ptr = np.concatenate([[0], np.cumsum([len(x) for x in tails])]).astype(np.int64)
    idx = np.concatenate(tails).astype(np.int64)
    orc = ko.OracleIndex(E, R, tr, va, te)
    lcg0 = rd.RefOpenKE.lcg_seeds(1)[0]                        # next_random[0] as randReset left it (nothing sampled since)
    got, used, _ = orc.corrupt_typed_words(ptr, idx, h, r, words, lcg0)
    assert used < len(words)
    assert np.array_equal(got, want), "oracle corrupt() != reference corrupt()"
    dup = sum(len(x) - len(np.unique(x)) for x in tails)
    assert dup == 0        # => the lists equal golden_type_constrain.npz's (sorted, unique), which the tests feed the oracle
    n_fb = int(sum(all(orc.find(int(a), int(t), int(b)) for t in tails[b]) for a, b in zip(h, r)))
    print(f"{len(h)} pairs, {used} rand() words consumed, {n_fb} corrupt_head fallbacks, repeated ids in the tail lists: {dup}")
    np.savez_compressed(os.path.join(HERE, "golden_corrupt.npz"), h=h.astype(np.uint16), r=r.astype(np.uint8), tails=want.astype(np.uint16),
                        words=words[:used].astype(np.int32), libc_seed=LIBC_SEED, lcg0=np.uint64(lcg0))


if __name__ == "__main__":
    main()
