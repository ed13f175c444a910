"""bench.py --workload zsl: the ZSL evaluation the reference's main.py really runs (ZSLmodule.eval, module/zsl_module.py:635-745)
at FB15K-237-ZS size: every test triple (17 596, 29 unseen relations) ranked among its candidate list (the relation's 1 000
rel2candidates minus the head's known tails, true tail first -- utils/gen_mode_candidates.py:15-39) by the Extractor + cosine
mean over 20 generated relation vectors.  Synthetic seeded Extractor weights, neighbour lists and relation vectors (the trained
ones cannot be obtained offline); ids and candidate sets are the bundled FB15K-237-ZS ones.  A "step" = one full evaluation:
entity halves + all (head, candidate) pairs + ranks + metrics.  Queries shard across ranks (test triples), weak scaling."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

D, MAX_NB, N_VEC = 200, 50, 20


def load(rank=0):
    import golden_util as gu
    z = gu.load("fb15k237_zs.npz")
    E, R = int(z["E"]), int(z["R"])
    h, r, t = (z[k].astype(np.int64) for k in ("test_h", "test_r", "test_t"))
    order = np.argsort(r, kind="stable")
    h, r, t = h[order], r[order], t[order]
    rel2cand = {int(rr): z["cand_ent"][i].astype(np.int64) for i, rr in enumerate(z["cand_rel"])}
    # candidate lists as utils/gen_mode_candidates.py:15-39 builds them: true tail first, then the relation's candidates that are
    # neither a known tail of (head, relation) nor the true tail
    known = {}
    for hh, rr, tt in zip(h.tolist(), r.tolist(), t.tolist()):
        known.setdefault((hh, rr), set()).add(tt)
    cands = []
    for hh, rr, tt in zip(h.tolist(), r.tolist(), t.tolist()):
        c = rel2cand[rr]
        keep = c[~np.isin(c, np.fromiter(known[(hh, rr)], np.int64)) & (c != tt)]
        cands.append(np.concatenate([[tt], keep]).astype(np.int64))
    rng = np.random.default_rng(192 + rank)
    n_symbols = E + R
    w = gu.seeded_extractor_weights(192, n_symbols, D)
    deg = rng.integers(1, MAX_NB + 1, E)
    conn = np.full((E, MAX_NB), n_symbols, np.int64)
    mask = np.arange(MAX_NB)[None, :] < deg[:, None]
    conn[mask] = rng.integers(0, E, int(mask.sum()))
    rels = np.unique(r)
    rel_slot = np.searchsorted(rels, r)
    rel_vecs = rng.standard_normal((len(rels), N_VEC, D)).astype(np.float32)
    return dict(E=E, R=R, w=w, conn=conn, deg=deg.astype(np.float32), heads=h, rel_slot=rel_slot, cands=cands, rel_vecs=rel_vecs)


def cpu_reference(wl, threads, sample):
    """the reference's loop: one Extractor forward per test triple on the gathered [C, 50] neighbour tensors, sklearn-style
    cosine mean, argsort (module/zsl_module.py:662-726), torch CPU on all host threads"""
    import torch
    from oracle import zsl_oracle as zo
    torch.set_num_threads(threads)
    conn, deg = wl["conn"], wl["deg"]
    t0 = time.perf_counter()
    for i in sample:
        hd, c = int(wl["heads"][i]), wl["cands"][i]
        left = np.full(len(c), hd)
        vecs = zo.extractor_query_vectors(wl["w"], np.stack([left, c], 1), conn[left], deg[left], conn[c], deg[c])
        s = zo.cosine_mean_scores(vecs, wl["rel_vecs"][wl["rel_slot"][i]])
        list(np.argsort(s))[::-1].index(0)
    return time.perf_counter() - t0


def main(args, rank, world, local):
    from bench import ClockSampler, measured_peaks, stdout_to_stderr
    wl = load(rank)
    T = len(wl["heads"])
    P = int(sum(len(c) for c in wl["cands"]))
    config = {"workload": "zsl", "desc": "ZSLmodule.eval on FB15K-237-ZS: 17 596 test triples x ~1 000 candidates, Extractor + cosine mean over 20 relation vectors",
              "E": wl["E"], "D": D, "triples_per_rank_per_step": T, "pairs_per_rank_per_step": P,
              "l2": "256 MiB buffer written between timed steps (L2 flush)", "sharding": "test triples sharded across ranks, weights replicated"}
    threads = os.cpu_count() or 1
    if args.impl == "reference":
        if rank != 0:
            return
        rng = np.random.default_rng(0)
        per_step = 24
        cpu_reference(wl, threads, rng.integers(0, T, 2))
        tot = 0.0
        for _ in range(args.steps):
            tot += cpu_reference(wl, threads, rng.integers(0, T, per_step))
        v = per_step * args.steps / tot
        args.emit({"impl": "reference", "metric": "ZSL eval test triples/sec", "value": v, "unit": "triples/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "triples/s", "cores": threads, "kind": "port",
                                           "sample": f"{per_step} test triples per step, one torch-CPU Extractor forward per triple (the reference's loop; its class is pinned to this restatement bit for bit)"},
                          "e2e": {"value": v, "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    import torch
    import torch.distributed as dist
    import mre_b200
    mre_b200.build()
    torch.cuda.set_device(local)
    dctx = mre_b200.dist.DistContext() if world > 1 else None
    dev = torch.device("cuda", local)
    ev = mre_b200.paper.ZSLEvaluator(wl["w"], wl["conn"], wl["deg"], np.arange(wl["E"]), device=local)
    ctx, L = ev.ctx, mre_b200._lib
    rk = mre_b200.engine.Ranker(ctx)
    # device-resident inputs for `value`; pinned host inputs for `e2e`
    ptr = np.concatenate([[0], np.cumsum([len(c) for c in wl["cands"]])]).astype(np.int64)
    flat = np.concatenate(wl["cands"])
    host = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (wl["heads"], wl["rel_slot"], ptr, flat)]
    d_head, d_rel, d_ptr, d_idx = (a.to(dev) for a in host)
    rv = torch.from_numpy(wl["rel_vecs"]).to(dev)
    counts = torch.empty((4, T), dtype=torch.int32, device=dev)
    counts_h = torch.empty((4, T), dtype=torch.int32).pin_memory()
    stream = lambda: torch.cuda.current_stream().cuda_stream

    def halves():
        L.check(L.lib().mre_zsl_entity_features(ctx._h, ev.model, ev.ent_symbol.data_ptr(), ev.conn.data_ptr(), ev.deg.data_ptr(),
                                                wl["E"], MAX_NB, ev.A.data_ptr(), ev.B.data_ptr(), stream()))

    def rank_all(dh, dr, dp, di):
        L.check(L.lib().mre_zsl_rank(ctx._h, ev.model, ev.A.data_ptr(), ev.B.data_ptr(), wl["E"], dh.data_ptr(), dr.data_ptr(), dp.data_ptr(),
                                     di.data_ptr(), T, P, rv.data_ptr(), rv.shape[0], N_VEC, None, counts.data_ptr(), stream()))

    def step_dev():
        halves()
        rank_all(d_head, d_rel, d_ptr, d_idx)
        return rk.metrics(counts, 1, "pessimistic")

    def step_e2e():
        halves()
        dh, dr, dp, di = (a.to(dev, non_blocking=True) for a in host)
        rank_all(dh, dr, dp, di)
        counts_h.copy_(counts, non_blocking=True)
        torch.cuda.synchronize()
        return counts_h

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    steps, warmup = args.steps, max(args.warmup, 3)
    for _ in range(warmup):
        step_dev()
    barrier()
    ctx.timing(True); ctx.timing_read()
    l0 = ctx.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with ClockSampler(local) as clocks:
        for a, b in evs:
            flush.zero_()
            a.record(); out = step_dev(); b.record()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        launches = ctx.launches - l0
        kern_ms, kern_n = ctx.timing_read()
        ctx.timing(False)
        for _ in range(2):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_e2e()
        barrier()
        e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([ms, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = t.tolist()
    sums, rr = out["sums"].clone(), out["rr"].clone()
    if dctx is not None:
        sums, rr = dctx.all_reduce_metrics(sums, rr)
    summ = mre_b200.engine.summarize(sums.cpu().numpy(), rr.cpu().numpy())
    # per pair ONE 2D -> D contraction is left (the hidden layer is split per entity); it runs as 3 x TF32 on the tensor cores
    flops = 2.0 * P * (2 * D * D)
    kms = kern_ms / max(kern_n, 1)
    achieved = flops / (kms * 1e-3) / 1e12
    tf32_probe = ctx.probe_tf32_peak() / 1e12
    mp = measured_peaks()
    if mp and mp.get("bf16_tflops"):
        peak = float(mp["bf16_tflops"]) / 2.0
        peak_src = f"half of MEASURED_PEAKS.json bf16_tflops (TF32 runs at half the BF16 rate; no TF32 figure in the file); in-process tcgen05 kind::tf32 probe: {tf32_probe:.0f} TFLOP/s"
    else:
        peak, peak_src = tf32_probe, "tcgen05 kind::tf32 dense MMA microbenchmark run in this process (mre_probe_tf32_peak); MEASURED_PEAKS.json absent"
    cpu_base = None
    if rank == 0 and not args.no_extra:
        rng = np.random.default_rng(0)
        cpu_reference(wl, threads, rng.integers(0, T, 2))
        n_s = 96
        s = cpu_reference(wl, threads, rng.integers(0, T, n_s))
        cpu_base = {"value": n_s / s, "unit": "triples/s", "cores": threads, "kind": "port",
                    "sample": f"{n_s} test triples of the same workload in {s:.1f} s: one torch-CPU Extractor forward per triple (the reference's loop)"}
    traffic, traffic_src = None, None
    try:        # DRAM bytes of the pair kernel per launch from the committed `ncu --set full` capture
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r2_traffic.json")) as f:
            tr = json.load(f).get("zsl")
        traffic, traffic_src = tr["dram_bytes_read"] + tr["dram_bytes_write"], tr["source"]
    except Exception:  # noqa: BLE001
        pass
    if rank == 0:
        args.emit({
            "metric": "ZSL eval test triples/sec", "value": world * T * steps / (ms * 1e-3), "unit": "triples/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "zsl_tc_kernel", "kernel_ms": kms, "launches_timed": kern_n, "peak_source": peak_src,
                         "pipe_frac": 3 * 208.0 / 200.0 * achieved / peak,
                         "algorithmic": "2 * pairs * 2D * D flops: the one per-pair contraction left after the per-entity split of the hidden layer; "
                                        "FP32-exact results need 3 TF32 products per FP32 product and N is padded 200 -> 208, so frac <= 0.32 by "
                                        "construction; pipe_frac = executed flops / peak.  LayerNorm / cosine epilogue and the operand gather not counted"},
            "cpu_baseline": cpu_base,
            "e2e": {"value": world * T * steps / e2e_s, "unit": "triples/s", "h2d_bytes_per_step": int(sum(a.numel() * 8 for a in host)),
                    "d2h_bytes_per_step": int(counts_h.numel() * 4), "api": "mre_zsl_entity_features + mre_zsl_rank: pinned host candidate lists in, int32 rank counts out"},
            "gpu_launches": int(launches), "clocks": clocks.summary(), "result": {"tail": summ[1]}})
