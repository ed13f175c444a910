"""Generates tests/golden/golden_siblings2.npz from the REFERENCE's own modules (BUILD CONTAINER ONLY): TransH, TransD and
Analogy predict (OpenKE/openke/module/model/{TransH,TransD,Analogy}.py) on seeded tables for 1-vs-all head and tail queries of
FB15K237 test triples -> the counts Test.h's compare loop gives those scores (oracle/kge_oracle.c, pinned to Base.so), the
1e-5 relative tie-band interval of every filtered count, and probe scores of the true entity and a few others."""
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
from oracle import kge_oracle as ko  # noqa: E402

N_QUERIES, D = 64, 96          # D = 96 keeps the fixture's tables cheap to rebuild in the tests; Analogy: 96 + 192


def build(name, E, R, wfn):
    from openke.module.model import Analogy, TransD, TransH
    if name == "transh":
        m = TransH(E, R, dim=D, p_norm=1, norm_flag=True)
        names = ["ent_embeddings", "rel_embeddings", "norm_vector"]
        shapes = [(E, D), (R, D), (R, D)]
    elif name == "transd":
        m = TransD(E, R, dim_e=D, dim_r=D, p_norm=1, norm_flag=True)
        names = ["ent_embeddings", "rel_embeddings", "ent_transfer", "rel_transfer"]
        shapes = [(E, D), (R, D), (E, D), (R, D)]
    else:
        m = Analogy(E, R, dim=D)
        names = ["ent_re_embeddings", "ent_im_embeddings", "rel_re_embeddings", "rel_im_embeddings", "ent_embeddings", "rel_embeddings"]
        shapes = [(E, D), (E, D), (R, D), (R, D), (E, 2 * D), (R, 2 * D)]
    tabs = wfn(gu.SEED + 7, shapes)
    for n, t in zip(names, tabs):
        getattr(m, n).weight.data.copy_(torch.from_numpy(t))
    return m


def main():
    z = gu.load("fb15k237_ids.npz")
    E, R = int(z["E"]), int(z["R"])
    splits = tuple(gu.split_cols(z, s) for s in ("train", "valid", "test"))
    ix = ko.OracleIndex(E, R, *splits)
    th, tt, tr = ix.test_triples()
    all_h, all_t, all_r = (np.concatenate([s[k] for s in splits]) for k in range(3))
    tails_of, heads_of = gu.group_lists(all_h, all_r, all_t), gu.group_lists(all_t, all_r, all_h)
    qidx = np.linspace(0, len(th) - 1, N_QUERIES).astype(np.int64)
    probe = np.random.default_rng(3).choice(E, 16, replace=False).astype(np.int64)
    out = {"qidx": qidx, "D": D, "probe": probe}
    ar = torch.arange(E)
    for wname, wfn in gu.WEIGHT_SETS.items():
        for name in ("transh", "transd", "analogy"):
            m = build(name, E, R, wfn)
            rows, probes = [], []
            for i in qidx.tolist():
                h, t, r = int(th[i]), int(tt[i]), int(tr[i])
                for side in (0, 1):
                    data = ({"batch_h": ar, "batch_t": torch.tensor([t]), "batch_r": torch.tensor([r]), "mode": "head_batch"} if side == 0
                            else {"batch_h": torch.tensor([h]), "batch_t": ar, "batch_r": torch.tensor([r]), "mode": "tail_batch"})
                    with torch.no_grad():
                        s = np.ascontiguousarray(m.predict(data), np.float32)
                    raw, filt = ix.rank_from_scores(s, side, h, t, r)
                    truth = h if side == 0 else t
                    known = heads_of.get((t, r), np.zeros(0, np.int64)) if side == 0 else tails_of.get((h, r), np.zeros(0, np.int64))
                    band = gu.TIE_BAND * max(abs(float(s[truth])), float(np.abs(s).mean()))
                    lo, hi = gu.band_counts(s, truth, np.asarray(known, np.int64), band)
                    assert lo <= filt <= hi
                    rows.append((raw, filt, lo, hi))
                    probes.append(np.concatenate([[s[truth]], s[probe]]))
            rows = np.asarray(rows, np.int64).reshape(len(qidx), 2, 4)
            key = f"{wname}_{name}"
            out[key + "_raw"], out[key + "_filt"] = rows[:, :, 0].astype(np.int32), rows[:, :, 1].astype(np.int32)
            out[key + "_lo"], out[key + "_hi"] = rows[:, :, 2].astype(np.int32), rows[:, :, 3].astype(np.int32)
            out[key + "_probe"] = np.asarray(probes, np.float32).reshape(len(qidx), 2, -1)
            print(key, "mean filt", rows[:, :, 1].mean(), "band-open", int((rows[:, :, 2] != rows[:, :, 3]).sum()))
    np.savez_compressed(os.path.join(HERE, "golden_siblings2.npz"), **out)
    print("golden_siblings2.npz", os.path.getsize(os.path.join(HERE, "golden_siblings2.npz")))


if __name__ == "__main__":
    main()
