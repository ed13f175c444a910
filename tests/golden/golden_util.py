"""Shared by tests/golden/make_golden.py (build container, reads /root/reference) and the tests
(any box, reads only the committed .npz fixtures)."""
import os

import numpy as np

GOLDEN_DIR = os.path.dirname(os.path.abspath(__file__))
SEED = 192  # the reference's default seed (args.py:8)
TIE_BAND = 1e-5  # north_star: ranks are exact for score gaps > 1e-5 relative


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name))


def xavier_tables(seed, shapes):
    """Seeded xavier-uniform float32 tables (the init OpenKE uses, TransE.py:21-22), drawn with numpy's
    PCG64 so the same bits come back on any box."""
    rng = np.random.default_rng(seed)
    out = []
    for rows, dim in shapes:
        a = np.sqrt(6.0 / (rows + dim))
        out.append(rng.uniform(-a, a, (rows, dim)).astype(np.float32))
    return out


def structured_tables(seed, shapes, rank=8):
    """Seeded low-rank-plus-noise float32 tables: scores spread over a wide range (as a trained model's
    do), so almost no candidate falls inside the 1e-5 tie band and exact rank equality is demanded."""
    rng = np.random.default_rng(seed + 1)
    out = []
    for rows, dim in shapes:
        z = rng.standard_normal((rows, rank))
        w = rng.standard_normal((rank, dim))
        x = np.zeros((rows, dim))
        for k in range(rank):  # explicit order: no BLAS, same bits on any CPU
            x += z[:, k:k + 1] * w[k:k + 1, :]
        x = x / np.sqrt(rank) + 0.1 * rng.standard_normal((rows, dim))
        out.append((x / np.sqrt(dim)).astype(np.float32))
    return out


WEIGHT_SETS = {"xavier": xavier_tables, "structured": structured_tables}


def split_cols(z, prefix):
    return tuple(z[f"{prefix}_{c}"].astype(np.int64) for c in "htr")


def group_lists(key_a, key_b, val):
    """dict (a, b) -> sorted unique array of val."""
    out = {}
    for a, b, v in zip(key_a.tolist(), key_b.tolist(), val.tolist()):
        out.setdefault((a, b), set()).add(v)
    return {k: np.fromiter(sorted(v), np.int64, len(v)) for k, v in out.items()}


def band_counts(scores, true_idx, known, band):
    """Filtered strict-rank interval implied by a float32 score vector under a tie band:
    lo = #{j unfiltered : s_j < s_true - band},  hi = #{j unfiltered : s_j <= s_true + band}."""
    s_true = scores[true_idx]
    mask = np.ones(len(scores), bool)
    mask[known] = False
    mask[true_idx] = False
    s = scores[mask]
    return int((s < s_true - band).sum()), int((s <= s_true + band).sum())
