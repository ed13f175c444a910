"""Sibling scorers and losses on the existing skeletons, against fixtures produced by the reference's own modules
(tests/golden/make_golden_siblings.py): SimplE.predict -> Base.so counts.  The losses are in tests/test_losses.py."""
import numpy as np
import pytest
import torch

import golden_util as gu


@pytest.mark.gpu
@pytest.mark.parametrize("wname", list(gu.WEIGHT_SETS))
def test_simple_ranks_vs_reference_golden(mre, fb15k237, wname):
    """SimplE ranks through the DistMult tcgen05 path: inside the reference's tie band, equal where the band is empty"""
    g = gu.load("golden_siblings.npz")
    eng = mre.engine
    E, R, D = fb15k237.E, fb15k237.R, int(g["D"])
    ent, rel, _, rel_inv = gu.WEIGHT_SETS[wname](gu.SEED, [(E, D), (R, D), (E, D), (R, D)])
    model = mre.openke.module.model.SimplE(E, R, dim=D)
    model.ent_embeddings.weight.data.copy_(torch.from_numpy(ent))
    model.rel_embeddings.weight.data.copy_(torch.from_numpy(rel))
    model.rel_inv_embeddings.weight.data.copy_(torch.from_numpy(rel_inv))
    model.cuda()
    ix = eng.KGIndex.from_arrays(E, R, fb15k237.train, fb15k237.valid, fb15k237.test).to_device(0)
    th, tt, tr = fb15k237.oracle.test_triples()
    qidx = g["qidx"]
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    q_h, q_t, q_r = np.repeat(th[qidx], 2), np.repeat(tt[qidx], 2), np.repeat(tr[qidx], 2)
    side = np.tile(np.array([0, 1], np.uint8), len(qidx))
    tabs = tuple(t.detach().contiguous() for t in model.tables())
    c = model.ranker().rank(model.scorer, tabs, dev(q_h), dev(q_t), dev(q_r), dev(side), index=ix).cpu().numpy()
    filt = c[2].reshape(-1, 2)
    lo, hi, ref = g[f"{wname}_simple_lo"], g[f"{wname}_simple_hi"], g[f"{wname}_simple_filt"]
    # head queries: the reference associates (h * r) * t, the kernel (r * t) . h -- both inside the 1e-5 relative band
    assert np.all(filt >= lo) and np.all(filt <= hi), (int(((filt < lo) | (filt > hi)).sum()),)
    exact = lo == hi
    assert np.array_equal(filt[exact], ref[exact])
    # predict parity for one query of each side
    for data in ({"batch_h": np.arange(E), "batch_t": tt[qidx[:1]], "batch_r": tr[qidx[:1]], "mode": "head_batch"},
                 {"batch_h": th[qidx[:1]], "batch_t": np.arange(E), "batch_r": tr[qidx[:1]], "mode": "tail_batch"}):
        s = model.predict({k: (torch.from_numpy(np.ascontiguousarray(v)).cuda() if k != "mode" else v) for k, v in data.items()})
        h = torch.from_numpy(ent[data["batch_h"]]); t = torch.from_numpy(ent[data["batch_t"]]); r = torch.from_numpy(rel[data["batch_r"]])
        want = -(torch.sum(h * r * t, -1)).numpy()
        assert np.allclose(s, want, rtol=2e-5, atol=1e-6)
