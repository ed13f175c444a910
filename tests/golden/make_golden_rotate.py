"""Generates tests/golden/golden_rotate.npz from the REFERENCE's own module (BUILD CONTAINER ONLY): RotatE.predict
(OpenKE/openke/module/model/RotatE.py:44-91) on seeded tables for 1-vs-all head and tail queries of FB15K237 test triples -> the
counts Test.h's compare loop gives those scores (oracle/kge_oracle.c, pinned to Base.so), the 1e-5 relative tie-band interval of
every filtered count, and probe scores of the true entity and a few others."""
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

N_QUERIES, D = 64, 50          # 50 complex dimensions: entity rows of 100 floats, a chunk of 16 + a ragged tail in the tile kernel


def build(name, E, R, wfn):
    from openke.module.model import RotatE
    m = RotatE(E, R, dim=D, margin=6.0, epsilon=2.0)
    # seeded tables scaled into RotatE's own initialisation ranges (RotatE.py:21-39): phases then cover (-pi, pi)
    ent, rel = wfn(gu.SEED + 11, [(E, 2 * D), (R, D)])
    ent = ent / np.abs(ent).max() * m.ent_embedding_range.item()
    rel = rel / np.abs(rel).max() * m.rel_embedding_range.item()
    m.ent_embeddings.weight.data.copy_(torch.from_numpy(ent.astype(np.float32)))
    m.rel_embeddings.weight.data.copy_(torch.from_numpy(rel.astype(np.float32)))
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
        for name in ("rotate",):
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
    np.savez_compressed(os.path.join(HERE, "golden_rotate.npz"), **out)
    print("golden_rotate.npz", os.path.getsize(os.path.join(HERE, "golden_rotate.npz")))


if __name__ == "__main__":
    main()
