"""TestDataLoader -- re-created from the C contract it wrapped:
  OpenKE/openke/base/Test.h:36-53      getHeadBatch / getTailBatch: 1-vs-all index arrays per test triple
  OpenKE/openke/config/Tester.py:72-82 set_sampling_mode('link'); iteration yields [data_head, data_tail]
  OpenKE/openke/base/Reader.h:227      test triples in (r, h, t) order
Iterating yields the reference's per-triple dict pairs (for callers that still want Model.predict per query); the
fused Tester never iterates -- it reads `queries()` and ranks every query in one pass.
"""
import numpy as np

from ... import engine


class TestDataSampler(object):
    def __init__(self, data_total, data_sampler):
        self.data_total = data_total
        self.data_sampler = data_sampler
        self.total = 0

    def __iter__(self):
        return self

    def __next__(self):
        self.total += 1
        if self.total > self.data_total:
            raise StopIteration()
        return self.data_sampler()

    def __len__(self):
        return self.data_total


class TestDataLoader(object):
    def __init__(self, in_path="./", sampling_mode="link", type_constrain=True, index=None):
        self.in_path = in_path
        self.sampling_mode = sampling_mode
        self.index = index if index is not None else engine.KGIndex.from_dir(in_path)   # loads type_constrain.txt if present
        self.type_constrain = type_constrain and self.index.has_type_constrain          # importTypeFiles, Reader.h:267-317
        self.relTotal, self.entTotal, self.testTotal = self.index.rel_tot, self.index.ent_tot, self.index.test_tot
        self.test_h, self.test_t, self.test_r = self.index.test_triples()
        self._ar = np.arange(self.entTotal, dtype=np.int64)
        self._cursor = 0

    def queries(self):
        """(q_h, q_t, q_r, q_side) of the whole link-prediction run in Tester order: head query then tail query per triple"""
        q_h, q_t, q_r = np.repeat(self.test_h, 2), np.repeat(self.test_t, 2), np.repeat(self.test_r, 2)
        side = np.tile(np.array([0, 1], np.uint8), self.testTotal)
        return q_h, q_t, q_r, side

    def type_groups(self, side, device, lo=0, hi=None):
        """CandidateGroups over the relation blocks of test triples [lo, hi) for the head (0) / tail (1) type lists"""
        hi = self.testTotal if hi is None else hi
        ptr, idx = self.index.type_constrain(side)
        r = self.test_r[lo:hi]
        qptr = np.searchsorted(r, np.arange(self.relTotal + 1), side="left")
        return engine.CandidateGroups(qptr, ptr, engine.torch.from_numpy(idx).to(device))

    def sampling_lp(self):
        i = self._cursor
        self._cursor += 1
        h, t, r = self.test_h[i:i + 1], self.test_t[i:i + 1], self.test_r[i:i + 1]
        return [
            {"batch_h": self._ar, "batch_t": t, "batch_r": r, "mode": "head_batch"},
            {"batch_h": h, "batch_t": self._ar, "batch_r": r, "mode": "tail_batch"},
        ]

    def get_ent_tot(self):
        return self.entTotal

    def get_rel_tot(self):
        return self.relTotal

    def get_triple_tot(self):
        return self.testTotal

    def set_sampling_mode(self, sampling_mode):
        self.sampling_mode = sampling_mode

    def __len__(self):
        return self.testTotal

    def __iter__(self):
        if self.sampling_mode != "link":
            raise NotImplementedError("triple classification (Test.h:396-422) is outside the hot path")
        self._cursor = 0
        return TestDataSampler(self.testTotal, self.sampling_lp)
