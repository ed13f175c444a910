"""Shared by the tests: fixtures from tests/golden and the oracle objects built from them."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import golden_util as gu  # noqa: E402
from oracle import kge_oracle as ko  # noqa: E402


class Dataset:
    pass


def load_fb15k237():
    z = gu.load("fb15k237_ids.npz")
    d = Dataset()
    d.E, d.R = int(z["E"]), int(z["R"])
    d.train, d.valid, d.test = (gu.split_cols(z, s) for s in ("train", "valid", "test"))
    d.oracle = ko.OracleIndex(d.E, d.R, d.train, d.valid, d.test)
    return d


def synthetic_graph(seed, E, R, n_train, n_valid, n_test):
    rng = np.random.default_rng(seed)
    def split(n):
        return (rng.integers(0, E, n), rng.integers(0, E, n), rng.integers(0, R, n))
    d = Dataset()
    d.E, d.R = E, R
    d.train, d.valid, d.test = split(n_train), split(n_valid), split(n_test)
    d.oracle = ko.OracleIndex(E, R, d.train, d.valid, d.test)
    return d


def oracle_counts(ds, score_fn, q_h, q_t, q_r, q_side):
    """(raw_lt, filt_lt) per query from the oracle: score_fn(side, h, t, r) -> float32[E]."""
    raw, filt = [], []
    for h, t, r, s in zip(q_h.tolist(), q_t.tolist(), q_r.tolist(), q_side.tolist()):
        con = score_fn(s, h, t, r)
        a, b = ds.oracle.rank_from_scores(con, s, h, t, r)
        raw.append(a); filt.append(b)
    return np.asarray(raw), np.asarray(filt)
