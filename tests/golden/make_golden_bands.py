"""Generates tests/golden/golden_bands.npz: the 1e-5 relative tie-band intervals of the REFERENCE's scores
for the two fixtures that only held counts -- golden_type_constrain.npz (type-constrained counts, Test.h:88-98,153-163) and
golden_siblings.npz (SimplE.predict, OpenKE/openke/module/model/SimplE.py:47-55) -- so their GPU tests can state the
north_star's bar (inside the band, equal where it is empty) instead of a looser one.  The scores come from the reference's own
modules (imported from /root/reference/OpenKE) and are checked to reproduce the committed counts.  BUILD CONTAINER ONLY."""
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

D = 200


def main():
    from openke.module.model import DistMult, SimplE, TransE
    z = gu.load("fb15k237_ids.npz")
    E, R = int(z["E"]), int(z["R"])
    splits = tuple(gu.split_cols(z, s) for s in ("train", "valid", "test"))
    ix = ko.OracleIndex(E, R, *splits)
    th, tt, tr = ix.test_triples()
    all_h, all_t, all_r = (np.concatenate([s[k] for s in splits]) for k in range(3))
    tails_of, heads_of = gu.group_lists(all_h, all_r, all_t), gu.group_lists(all_t, all_r, all_h)
    tc = gu.load("golden_type_constrain.npz")
    sib = gu.load("golden_siblings.npz")
    hp, tp = tc["head_ptr"], tc["tail_ptr"]
    hi_, ti_ = tc["head_idx"].astype(np.int64), tc["tail_idx"].astype(np.int64)
    ar = torch.arange(E)
    out = {}

    def predict(m, side, h, t, r):
        data = ({"batch_h": ar, "batch_t": torch.tensor([t]), "batch_r": torch.tensor([r]), "mode": "head_batch"} if side == 0
                else {"batch_h": torch.tensor([h]), "batch_t": ar, "batch_r": torch.tensor([r]), "mode": "tail_batch"})
        with torch.no_grad():
            return m.predict(data)

    for wname, wfn in gu.WEIGHT_SETS.items():
        ent, rel, _, rel_inv = wfn(gu.SEED, [(E, D), (R, D), (E, D), (R, D)])
        models = {"transe_l1_norm": TransE(E, R, dim=D, p_norm=1, norm_flag=True), "distmult": DistMult(E, R, dim=D), "simple": SimplE(E, R, dim=D)}
        for m in models.values():
            m.ent_embeddings.weight.data.copy_(torch.from_numpy(ent)); m.rel_embeddings.weight.data.copy_(torch.from_numpy(rel))
        models["simple"].rel_inv_embeddings.weight.data.copy_(torch.from_numpy(rel_inv))
        # ---- type-constrained counts
        for name in ("transe_l1_norm", "distmult"):
            rows = []
            for k, i in enumerate(tc["qidx"].tolist()):
                h, t, r = int(th[i]), int(tt[i]), int(tr[i])
                for side in (0, 1):
                    s = predict(models[name], side, h, t, r)
                    lst = hi_[hp[r]:hp[r + 1]] if side == 0 else ti_[tp[r]:tp[r + 1]]
                    raw, filt = ix.rank_from_scores_constrained(s, side, h, t, r, lst)
                    assert (raw, filt) == (tc[f"{wname}_{name}_raw"][k, side], tc[f"{wname}_{name}_filt"][k, side])
                    truth = h if side == 0 else t
                    known = heads_of[(t, r)] if side == 0 else tails_of[(h, r)]
                    band = gu.TIE_BAND * max(abs(float(s[truth])), float(np.abs(s).mean()))
                    cand = np.setdiff1d(lst, np.concatenate([known, [truth]]))
                    lo, hi = int((s[cand] < s[truth] - band).sum()), int((s[cand] <= s[truth] + band).sum())
                    assert lo <= filt <= hi
                    rows.append((lo, hi))
            rows = np.asarray(rows, np.int32).reshape(-1, 2, 2)
            out[f"tc_{wname}_{name}_lo"], out[f"tc_{wname}_{name}_hi"] = rows[:, :, 0], rows[:, :, 1]
            print("type-constrained", wname, name, "band-open queries:", int((rows[:, :, 0] != rows[:, :, 1]).sum()), "/", rows.shape[0] * 2)
        # ---- SimplE: the committed intervals used band = 1e-5 * max(|s_true|, mean|s|) already; re-derive and compare
        rows = []
        for k, i in enumerate(sib["qidx"].tolist()):
            h, t, r = int(th[i]), int(tt[i]), int(tr[i])
            for side in (0, 1):
                s = predict(models["simple"], side, h, t, r)
                raw, filt = ix.rank_from_scores(s, side, h, t, r)
                assert filt == sib[f"{wname}_simple_filt"][k, side]
                truth = h if side == 0 else t
                known = heads_of[(t, r)] if side == 0 else tails_of[(h, r)]
                band = gu.TIE_BAND * max(abs(float(s[truth])), float(np.abs(s).mean()))
                rows.append(gu.band_counts(s, truth, known, band))
        rows = np.asarray(rows, np.int32).reshape(-1, 2, 2)
        out[f"simple_{wname}_lo"], out[f"simple_{wname}_hi"] = rows[:, :, 0], rows[:, :, 1]
        print("SimplE", wname, "band-open queries:", int((rows[:, :, 0] != rows[:, :, 1]).sum()), "/", rows.shape[0] * 2,
              "| committed interval identical:", bool(np.array_equal(rows[:, :, 0], sib[f"{wname}_simple_lo"]) and np.array_equal(rows[:, :, 1], sib[f"{wname}_simple_hi"])))
    np.savez_compressed(os.path.join(HERE, "golden_bands.npz"), **out)
    print("golden_bands.npz", os.path.getsize(os.path.join(HERE, "golden_bands.npz")))


if __name__ == "__main__":
    main()
