"""Tester (OpenKE/openke/config/Tester.py:17-91): same constructor and the same (mrr, mr, hit10, hit3, hit1) tuple from
run_link_prediction, computed in ONE fused pass over all 2 x testTotal queries (mre_rank + mre_metrics) instead of one
Model.predict + host float[E] copy + testHead/testTail call per query."""
import numpy as np
import torch

from ... import engine


class Tester(object):
    def __init__(self, model=None, data_loader=None, use_gpu=True):
        self.model = model
        self.data_loader = data_loader
        self.use_gpu = use_gpu
        if self.use_gpu and self.model is not None:
            self.model.cuda()
        self.last = None     # per-side summary of the most recent run (also Hits@5, the paper's metric set)

    def set_model(self, model):
        self.model = model

    def set_data_loader(self, data_loader):
        self.data_loader = data_loader

    def set_use_gpu(self, use_gpu):
        self.use_gpu = use_gpu
        if self.use_gpu and self.model is not None:
            self.model.cuda()

    def to_var(self, x, use_gpu):
        t = torch.from_numpy(np.ascontiguousarray(x))
        return t.cuda() if use_gpu else t

    def test_one_step(self, data):                                              # Tester.py:62-68
        return self.model.predict({
            "batch_h": self.to_var(data["batch_h"], self.use_gpu),
            "batch_t": self.to_var(data["batch_t"], self.use_gpu),
            "batch_r": self.to_var(data["batch_r"], self.use_gpu),
            "mode": data["mode"],
        })

    def rank_counts(self, dist=None):
        """device int32 counts [4, Q] of this rank's shard of the link-prediction queries (+ the shard's side array)"""
        if not self.use_gpu:
            raise engine.L.MreError("mre_b200 ranks on a B200 only (use_gpu=False has no fallback)")
        dl, m = self.data_loader, self.model
        dev = m.device()
        if dl.index.device is None:
            dl.index.to_device(dev.index or 0)
        q_h, q_t, q_r, side = dl.queries()
        if dist is not None:
            lo, hi = dist.shard(len(q_h))
            q_h, q_t, q_r, side = q_h[lo:hi], q_t[lo:hi], q_r[lo:hi], side[lo:hi]
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        side_d = to(side)
        if hasattr(m, "rank_queries"):          # per-relation projected tables (TransH, TransD): model/_projected.py
            return m.rank_queries(q_h, q_t, q_r, side, dl.index), side_d
        tabs = tuple(t.detach().contiguous() for t in m.tables())
        counts = m.ranker().rank(m.scorer, tabs, to(q_h), to(q_t), to(q_r), side_d, index=dl.index, **m.rank_kwargs())
        return counts, side_d

    def rank_counts_constrained(self, dist=None):
        """type-constrained counts (Test.h:88-98,153-163): one grouped job per side, a candidate group per relation over
        the relation's head / tail list of type_constrain.txt; the true entity's score is the threshold even when the
        true entity is not in the list, exactly as `minimal = con[h]` is taken before the constraint test"""
        if not self.use_gpu:
            raise engine.L.MreError("mre_b200 ranks on a B200 only (use_gpu=False has no fallback)")
        dl, m = self.data_loader, self.model
        if not dl.index.has_type_constrain:
            raise engine.L.MreError("type_constrain=True needs type_constrain.txt (importTypeFiles, Reader.h:267-317)")
        dev = m.device()
        if dl.index.device is None:
            dl.index.to_device(dev.index or 0)
        lo, hi = (0, dl.testTotal) if dist is None else dist.shard(dl.testTotal)
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        q_h, q_t, q_r = to(dl.test_h[lo:hi]), to(dl.test_t[lo:hi]), to(dl.test_r[lo:hi])
        tabs = tuple(t.detach().contiguous() for t in m.tables())
        out = []
        for side in (0, 1):
            groups = dl.type_groups(side, dev, lo, hi)
            out.append(m.ranker().rank(m.scorer, tabs, q_h, q_t, q_r, side, index=dl.index, groups=groups, **m.rank_kwargs()))
        return out

    def run_link_prediction(self, type_constrain=False, dist=None):              # Tester.py:70-91
        self.data_loader.set_sampling_mode("link")
        if type_constrain:
            # the reference fills both metric sets in this mode and returns the constrained one (Test.h:352-390)
            sums = torch.zeros((2, 8), dtype=torch.int64, device=self.model.device())
            rr = torch.zeros(2, dtype=torch.float64, device=self.model.device())
            for side, counts in enumerate(self.rank_counts_constrained(dist)):
                o = self.model.ranker().metrics(counts, side, "strict")
                sums += o["sums"]
                rr += o["rr"]
        else:
            counts, side_d = self.rank_counts(dist)
            out = self.model.ranker().metrics(counts, side_d, "strict")
            sums, rr = out["sums"], out["rr"]
        if dist is not None:
            sums, _ = dist.all_reduce_metrics(sums)
        sums = sums.cpu().numpy()
        rr = sums[:, 6].astype(np.float64) / engine.RR_FIXED_ONE    # integer fixed-point sums: the same bits for any sharding
        self.last = engine.summarize(sums)
        # test_link_prediction (Test.h:232-277): each side divided by testTotal, then (head + tail) / 2 of the filtered values
        T = float(self.data_loader.get_triple_tot())
        mrr = (rr[0] + rr[1]) / T / 2
        mr = (sums[0][1] + sums[1][1]) / T / 2
        hit10 = (sums[0][5] + sums[1][5]) / T / 2
        hit3 = (sums[0][3] + sums[1][3]) / T / 2
        hit1 = (sums[0][2] + sums[1][2]) / T / 2
        print(hit10)
        return mrr, mr, hit10, hit3, hit1
