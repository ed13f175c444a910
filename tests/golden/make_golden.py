"""Generate the committed fixtures under tests/golden/.  BUILD CONTAINER ONLY: reads /root/reference
(datasets, the reference's OpenKE PyTorch models) and oracle/_ref/Base.so (the compiled reference library).

    python tests/golden/make_golden.py

Fixtures written (all small, integer ids re-encoded; no reference source is copied):
  fb15k237_ids.npz     OpenKE/benchmarks/FB15K237/{train,valid,test}2id.txt as uint16/uint8 columns
  fb15k237_zs.npz      origin_data/FB15K-237-ZS test triples + rel2candidates as ids
  db15k_zs.npz         origin_data/DB15K-ZS test triples as ids
  golden_fb15k237.npz  per-query results of the REAL reference (torch-CPU Model.predict of the reference's
                       own modules -> Base.so testHead/testTail) on seeded weights, for a spread of test
                       triples: raw/filtered counts, s_true, tie-band interval, score probes, metric tuples
  golden_sampler.npz   Base.so `sampling` output with its LCG pinned (srand(1)), 2 threads
  golden_type_constrain.npz  FB15K237's type_constrain.txt sets + Base.so's type-constrained metric tuples (Test.h:88-98)
While generating, this script ASSERTS that the oracle restatements agree with the reference:
oracle/openke_torch.py bit-identical to the reference modules; oracle/kge_oracle.c counts -> the same metric
tuple as Base.so; orc_sample_lcg bit-identical to Base.so's sampler.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True
REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "OpenKE"))

import golden_util as gu  # noqa: E402
from oracle import kge_oracle as ko, openke_torch as ot, ref_driver as rd  # noqa: E402

N_QUERIES = 192
PROBE = 64


def read2id(path):
    a = np.loadtxt(path, skiprows=1, dtype=np.int64)
    return a[:, 0].copy(), a[:, 1].copy(), a[:, 2].copy()


def fb15k237_ids():
    p = os.path.join(REF, "OpenKE/benchmarks/FB15K237")
    E = int(open(os.path.join(p, "entity2id.txt")).readline())
    R = int(open(os.path.join(p, "relation2id.txt")).readline())
    out = {"E": E, "R": R}
    for split in ("train", "valid", "test"):
        h, t, r = read2id(os.path.join(p, f"{split}2id.txt"))
        out[f"{split}_h"], out[f"{split}_t"], out[f"{split}_r"] = h.astype(np.uint16), t.astype(np.uint16), r.astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "fb15k237_ids.npz"), **out)
    return out


def zs_ids(name, out_name, with_cands):
    p = os.path.join(REF, "origin_data", name)
    e_id = json.load(open(os.path.join(p, "entity2ids_zsl.json")))
    r_id = json.load(open(os.path.join(p, "relation2ids.json")))
    task = json.load(open(os.path.join(p, "test_tasks_zsl.json")))
    h, r, t = [], [], []
    for rel in task.keys():  # module/utils.py:194-207 (load_appendix_data) iteration order
        for head, rel_, tail in task[rel]:
            h.append(e_id[head]); r.append(r_id[rel_]); t.append(e_id[tail])
    out = {"E": len(e_id), "R": len(r_id), "test_h": np.asarray(h, np.int32), "test_r": np.asarray(r, np.int32),
           "test_t": np.asarray(t, np.int32)}
    if with_cands:
        c = json.load(open(os.path.join(p, "rel2candidates_all.json")))
        rels = [rel for rel in task.keys()]
        out["cand_rel"] = np.asarray([r_id[x] for x in rels], np.int32)
        out["cand_ent"] = np.asarray([[e_id[e] for e in c[x]] for x in rels], np.int32)
    np.savez_compressed(os.path.join(HERE, out_name), **out)
    print(out_name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


def golden_for_weights(wname, wfn, out, ref, ix, test, qidx, probe, heads_of, tails_of, E, R, D, classes):
    TransE, DistMult, ComplEx = classes
    th, tt, trr = test
    ent, rel, ent_im, rel_im = wfn(gu.SEED, [(E, D), (R, D), (E, D), (R, D)])
    configs = {
        "transe_l1_norm": ("transe", dict(p_norm=1, norm_flag=True)),
        "transe_l2_norm": ("transe", dict(p_norm=2, norm_flag=True)),
        "transe_l1_raw": ("transe", dict(p_norm=1, norm_flag=False)),
        "distmult": ("distmult", {}),
        "complex": ("complex", {}),
    }
    for name, (kind, kw) in configs.items():
        if kind == "transe":
            m = TransE(E, R, dim=D, **kw)
            m.ent_embeddings.weight.data.copy_(torch.from_numpy(ent)); m.rel_embeddings.weight.data.copy_(torch.from_numpy(rel))
            tables = (torch.from_numpy(ent), torch.from_numpy(rel))
        elif kind == "distmult":
            m = DistMult(E, R, dim=D)
            m.ent_embeddings.weight.data.copy_(torch.from_numpy(ent)); m.rel_embeddings.weight.data.copy_(torch.from_numpy(rel))
            tables = (torch.from_numpy(ent), torch.from_numpy(rel))
        else:
            m = ComplEx(E, R, dim=D)
            m.ent_re_embeddings.weight.data.copy_(torch.from_numpy(ent)); m.ent_im_embeddings.weight.data.copy_(torch.from_numpy(ent_im))
            m.rel_re_embeddings.weight.data.copy_(torch.from_numpy(rel)); m.rel_im_embeddings.weight.data.copy_(torch.from_numpy(rel_im))
            tables = tuple(torch.from_numpy(x) for x in (ent, ent_im, rel, rel_im))
        ref.L.initTest()
        acc = ko.MetricAccumulator()
        cols = {k: [] for k in ("raw", "filt", "s_true", "lo", "hi", "lo3", "hi3", "band", "probe_scores")}
        ar = np.arange(E, dtype=np.int64)
        for i in qidx.tolist():
            h, t, r = int(th[i]), int(tt[i]), int(trr[i])
            for side in (0, 1):
                if side == 0:
                    data = {"batch_h": ar, "batch_t": np.array([t]), "batch_r": np.array([r]), "mode": "head_batch"}
                else:
                    data = {"batch_h": np.array([h]), "batch_t": ar, "batch_r": np.array([r]), "mode": "tail_batch"}
                tdata = {k: (torch.from_numpy(v) if k != "mode" else v) for k, v in data.items()}
                with torch.no_grad():
                    s = m.predict(tdata)
                    s2 = ot.predict(kind, tables, tdata, **kw).numpy()
                assert s.dtype == np.float32 and np.array_equal(s, s2), f"openke_torch != reference for {name}"
                (ref.test_head if side == 0 else ref.test_tail)(s, i)
                raw, filt = ix.rank_from_scores(s, side, h, t, r)
                acc.add(side, raw, filt)
                truth = h if side == 0 else t
                known = heads_of[(t, r)] if side == 0 else tails_of[(h, r)]
                band = gu.TIE_BAND * max(abs(float(s[truth])), float(np.abs(s).mean()))
                lo, hi = gu.band_counts(s, truth, known, band)
                assert lo <= filt <= hi
                # 3x band: the interval for arithmetic that is not bit-matched to torch's (the tensor-core path); the
                # reference's own FP32 summation error reaches ~0.5 band on the structured tables
                lo3, hi3 = gu.band_counts(s, truth, known, 3 * band)
                for k, v in (("raw", raw), ("filt", filt), ("s_true", s[truth]), ("lo", lo), ("hi", hi), ("lo3", lo3), ("hi3", hi3), ("band", band),
                             ("probe_scores", s[probe].copy())):
                    cols[k].append(v)
        tup_ref = ref.finish()
        tup_orc = acc.final(ix.test_total)
        assert tup_ref == tup_orc, (name, tup_ref, tup_orc)
        n = len(qidx)
        out[f"{wname}_{name}_raw"] = np.asarray(cols["raw"], np.int32).reshape(n, 2)
        out[f"{wname}_{name}_filt"] = np.asarray(cols["filt"], np.int32).reshape(n, 2)
        out[f"{wname}_{name}_lo"] = np.asarray(cols["lo"], np.int32).reshape(n, 2)
        out[f"{wname}_{name}_hi"] = np.asarray(cols["hi"], np.int32).reshape(n, 2)
        out[f"{wname}_{name}_lo3"] = np.asarray(cols["lo3"], np.int32).reshape(n, 2)
        out[f"{wname}_{name}_hi3"] = np.asarray(cols["hi3"], np.int32).reshape(n, 2)
        out[f"{wname}_{name}_s_true"] = np.asarray(cols["s_true"], np.float32).reshape(n, 2)
        out[f"{wname}_{name}_band"] = np.asarray(cols["band"], np.float32).reshape(n, 2)
        out[f"{wname}_{name}_probe_scores"] = np.asarray(cols["probe_scores"], np.float32).reshape(n, 2, PROBE)
        out[f"{wname}_{name}_tuple"] = np.asarray(tup_ref, np.float32)
        amb = int((out[f"{wname}_{name}_lo"] != out[f"{wname}_{name}_hi"]).sum())
        print(wname, name, "tuple", tup_ref, "queries with a non-empty tie band:", amb, "/", 2 * n)


def main():
    ids = fb15k237_ids()
    zs_ids("FB15K-237-ZS", "fb15k237_zs.npz", True)
    zs_ids("DB15K-ZS", "db15k_zs.npz", False)

    E, R, D = int(ids["E"]), int(ids["R"]), 200
    tr, va, te = (tuple(ids[f"{s}_{c}"].astype(np.int64) for c in "htr") for s in ("train", "valid", "test"))
    ix = ko.OracleIndex(E, R, tr, va, te)
    th, tt, trr = ix.test_triples()

    # ---- sampler golden + LCG pin ----
    bench_dir = os.path.join(REF, "OpenKE/benchmarks/FB15K237/")
    ref = rd.RefOpenKE(bench_dir, threads=2, bern=1)
    seeds = rd.RefOpenKE.lcg_seeds(2)
    samp = {"seeds": np.asarray(seeds, np.uint64), "B": 256, "neg": 5}
    for step in range(2):
        h, t, r, y = ref.sampling(256, 5, 0)
        oh, ot_, orr, oy = ix.sample_lcg(seeds, 256, 5, 0, 1)
        assert np.array_equal(h, oh) and np.array_equal(t, ot_) and np.array_equal(r, orr) and np.array_equal(y, oy)
        samp[f"h{step}"], samp[f"t{step}"], samp[f"r{step}"] = h.astype(np.uint16), t.astype(np.uint16), r.astype(np.uint8)
    lm, rm = ix.means()
    samp["left_mean"], samp["right_mean"] = lm, rm
    np.savez_compressed(os.path.join(HERE, "golden_sampler.npz"), **samp)

    # ---- ranking goldens from the real reference models ----
    from openke.module.model import TransE, DistMult, ComplEx

    all_h = np.concatenate([tr[0], va[0], te[0]]); all_t = np.concatenate([tr[1], va[1], te[1]])
    all_r = np.concatenate([tr[2], va[2], te[2]])
    tails_of = gu.group_lists(all_h, all_r, all_t)
    heads_of = gu.group_lists(all_t, all_r, all_h)

    qidx = np.linspace(0, ix.test_total - 1, N_QUERIES).astype(np.int64)
    probe = np.random.default_rng(7).choice(E, PROBE, replace=False).astype(np.int64)
    probe.sort()
    out = {"qidx": qidx, "probe": probe, "D": D}
    ref.load_test(type_files=True)

    for wname, wfn in gu.WEIGHT_SETS.items():
        golden_for_weights(wname, wfn, out, ref, ix, (th, tt, trr), qidx, probe, heads_of, tails_of, E, R, D,
                           (TransE, DistMult, ComplEx))
    np.savez_compressed(os.path.join(HERE, "golden_fb15k237.npz"), **out)

    # ---- type-constrained ranking (Test.h:88-98) on the real type_constrain.txt: metric tuples of Base.so + the type sets
    heads, tails = rd.read_type_constrain(bench_dir)
    tc = {"R": R}
    tc["head_ptr"] = np.concatenate([[0], np.cumsum([len(heads[r]) for r in range(R)])]).astype(np.int64)
    tc["tail_ptr"] = np.concatenate([[0], np.cumsum([len(tails[r]) for r in range(R)])]).astype(np.int64)
    tc["head_idx"] = np.concatenate([heads[r] for r in range(R)]).astype(np.uint16)
    tc["tail_idx"] = np.concatenate([tails[r] for r in range(R)]).astype(np.uint16)
    tc["qidx"] = qidx
    ar = np.arange(E, dtype=np.int64)
    for wname, wfn in gu.WEIGHT_SETS.items():
        ent, rel, ent_im, rel_im = wfn(gu.SEED, [(E, D), (R, D), (E, D), (R, D)])
        for name, kind, kw in (("transe_l1_norm", "transe", dict(p_norm=1, norm_flag=True)), ("distmult", "distmult", {})):
            tables = (torch.from_numpy(ent), torch.from_numpy(rel))
            ref.L.initTest()
            acc = ko.MetricAccumulator()
            raws, filts = [], []
            for i in qidx.tolist():
                h, t, r = int(th[i]), int(tt[i]), int(trr[i])
                for side in (0, 1):
                    data = ({"batch_h": ar, "batch_t": np.array([t]), "batch_r": np.array([r]), "mode": "head_batch"} if side == 0
                            else {"batch_h": np.array([h]), "batch_t": ar, "batch_r": np.array([r]), "mode": "tail_batch"})
                    tdata = {k: (torch.from_numpy(v) if k != "mode" else v) for k, v in data.items()}
                    with torch.no_grad():
                        s = ot.predict(kind, tables, tdata, **kw).numpy()
                    (ref.test_head if side == 0 else ref.test_tail)(s, i, 1)
                    raw, filt = ix.rank_from_scores_constrained(s, side, h, t, r, heads[r] if side == 0 else tails[r])
                    acc.add(side, raw, filt)
                    raws.append(raw); filts.append(filt)
            tup_ref = ref.finish(1)
            tup_orc = acc.final(ix.test_total)
            assert tup_ref == tup_orc, (wname, name, tup_ref, tup_orc)
            tc[f"{wname}_{name}_tuple"] = np.asarray(tup_ref, np.float32)
            tc[f"{wname}_{name}_raw"] = np.asarray(raws, np.int32).reshape(-1, 2)
            tc[f"{wname}_{name}_filt"] = np.asarray(filts, np.int32).reshape(-1, 2)
            print("type-constrained", wname, name, tup_ref)
    np.savez_compressed(os.path.join(HERE, "golden_type_constrain.npz"), **tc)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
