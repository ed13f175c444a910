"""Generates tests/golden/golden_siblings.npz from the REFERENCE's own modules (BUILD CONTAINER ONLY):
  * SimplE.predict (OpenKE/openke/module/model/SimplE.py:47-55) on seeded tables -> Base.so testHead/testTail: raw / filtered
    counts, s_true, tie-band interval per query, and the metric tuple;
  * SigmoidLoss, SoftplusLoss, MarginLoss (plain and self-adversarial; OpenKE/openke/module/loss/*.py) on seeded score blocks.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/OpenKE")

import golden_util as gu  # noqa: E402
from oracle import kge_oracle as ko, ref_driver as rd  # noqa: E402

N_QUERIES, D = 96, 200


def main():
    from openke.module.model import SimplE
    from openke.module.loss import MarginLoss, SigmoidLoss, SoftplusLoss
    z = gu.load("fb15k237_ids.npz")
    E, R = int(z["E"]), int(z["R"])
    splits = tuple(gu.split_cols(z, s) for s in ("train", "valid", "test"))
    import tempfile
    d = rd.write_benchmark_dir(tempfile.mkdtemp(prefix="mre_sib_"), E, R, *splits)
    ref = rd.RefOpenKE(d + "/", threads=2)
    ref.load_test()
    ix = ko.OracleIndex(E, R, *splits)
    th, tt, tr = ix.test_triples()
    qidx = np.linspace(0, len(th) - 1, N_QUERIES).astype(np.int64)
    out = {"qidx": qidx, "D": D}
    ar = torch.arange(E)
    for wname, wfn in gu.WEIGHT_SETS.items():
        ent, rel, _, rel_inv = wfn(gu.SEED, [(E, D), (R, D), (E, D), (R, D)])
        m = SimplE(E, R, dim=D)
        m.ent_embeddings.weight.data.copy_(torch.from_numpy(ent))
        m.rel_embeddings.weight.data.copy_(torch.from_numpy(rel))
        m.rel_inv_embeddings.weight.data.copy_(torch.from_numpy(rel_inv))
        ref.L.initTest()
        rows = []
        for i in qidx.tolist():
            h, t, r = int(th[i]), int(tt[i]), int(tr[i])
            for side in (0, 1):
                data = ({"batch_h": ar, "batch_t": torch.tensor([t]), "batch_r": torch.tensor([r]), "mode": "head_batch"} if side == 0
                        else {"batch_h": torch.tensor([h]), "batch_t": ar, "batch_r": torch.tensor([r]), "mode": "tail_batch"})
                s = m.predict(data)
                (ref.test_head if side == 0 else ref.test_tail)(s, i)
                raw, filt = ix.rank_from_scores(s, side, h, t, r)
                truth = h if side == 0 else t
                known = [j for j in range(E) if j != truth and s[j] <= s[truth] + 1e-3 and
                         (ix.find(j, t, r) if side == 0 else ix.find(h, j, r))]
                band = gu.TIE_BAND * max(abs(float(s[truth])), float(np.abs(s).mean()))
                lo, hi = gu.band_counts(s, truth, np.asarray(known, np.int64), band)
                rows.append((raw, filt, lo, hi))
        rows = np.asarray(rows, np.int64).reshape(len(qidx), 2, 4)
        out[f"{wname}_simple_raw"], out[f"{wname}_simple_filt"] = rows[:, :, 0].astype(np.int32), rows[:, :, 1].astype(np.int32)
        out[f"{wname}_simple_lo"], out[f"{wname}_simple_hi"] = rows[:, :, 2].astype(np.int32), rows[:, :, 3].astype(np.int32)
        out[f"{wname}_simple_tuple"] = np.asarray(ref.finish(), np.float32)
        print(wname, "SimplE tuple", out[f"{wname}_simple_tuple"], "band-open queries", int((rows[:, :, 2] != rows[:, :, 3]).sum()))
    # losses on seeded score blocks
    rng = np.random.default_rng(gu.SEED)
    p = torch.from_numpy(rng.standard_normal((64, 1)).astype(np.float32) * 3)
    n = torch.from_numpy(rng.standard_normal((64, 25)).astype(np.float32) * 3)
    out["loss_p"], out["loss_n"] = p.numpy(), n.numpy()
    for name, cls, kw in (("margin", MarginLoss, dict(margin=5.0)), ("margin_adv", MarginLoss, dict(adv_temperature=1.0, margin=6.0)),
                          ("sigmoid", SigmoidLoss, {}), ("sigmoid_adv", SigmoidLoss, dict(adv_temperature=2.0)),
                          ("softplus", SoftplusLoss, {}), ("softplus_adv", SoftplusLoss, dict(adv_temperature=0.5))):
        out["loss_" + name] = np.asarray(cls(**kw)(p, n).detach().numpy(), np.float32).reshape(-1)
    np.savez_compressed(os.path.join(HERE, "golden_siblings.npz"), **out)
    print("golden_siblings.npz", os.path.getsize(os.path.join(HERE, "golden_siblings.npz")))


if __name__ == "__main__":
    main()
