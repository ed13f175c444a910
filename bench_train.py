"""bench.py --workload train: BASELINE configs[3], the OpenKE TransE training step on FB15K237 (B = 4096 positives x 25
Bernoulli negatives, margin 5, L1, normalised, SGD lr 1.0; OpenKE/examples/train_transe_FB15K237.py:9-39), data-parallel:
every rank samples its own Philox sub-stream, runs the fused forward+backward, all-reduces the two dense gradient
tables over NCCL and applies the replicated SGD update.  A "step" = sample + margin step + all-reduce + update."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def cpu_reference_step(w, threads, steps):
    """the reference's CPU step: Base.so `sampling` on all host threads + the reference tensor expressions with autograd"""
    import tempfile
    import torch
    from oracle import openke_torch as ot, ref_driver as rd, kge_oracle as ko
    torch.set_num_threads(threads)
    ent, rel = torch.from_numpy(w["ent"]), torch.from_numpy(w["rel"])
    kind = "port"
    if rd.available():
        d = tempfile.mkdtemp(prefix="mre_ref_train_")
        rd.write_benchmark_dir(d, w["E"], w["R"], w["train"], w["valid"], w["test"])
        ref = rd.RefOpenKE(d + "/", threads=threads, bern=1)
        draw = lambda: ref.sampling(w["B"], w["neg"])
        kind = "reference"
    else:
        ix = ko.OracleIndex(w["E"], w["R"], w["train"], w["valid"], w["test"])
        draw = lambda: ix.sample_philox(192, 0, w["B"], w["neg"])
    t0 = time.perf_counter()
    for _ in range(steps):
        h, t, r, y = draw()
        ot.transe_train_step(ent, rel, torch.from_numpy(h), torch.from_numpy(t), torch.from_numpy(r), w["B"], 5.0, 1, True)
    return (time.perf_counter() - t0) / steps, kind


def main(args, rank, world, local):
    import golden_util as gu
    from bench import ClockSampler, stdout_to_stderr
    z = gu.load("fb15k237_ids.npz")
    E, R, D, B, neg = int(z["E"]), int(z["R"]), 200, 4096, 25
    splits = tuple(gu.split_cols(z, s) for s in ("train", "valid", "test"))
    ent, rel = gu.xavier_tables(192, [(E, D), (R, D)])
    w = {"E": E, "R": R, "B": B, "neg": neg, "train": splits[0], "valid": splits[1], "test": splits[2], "ent": ent, "rel": rel}
    n = B * (1 + neg)
    config = {"workload": "train", "desc": "OpenKE TransE step on FB15K237: B=4096 x 25 Bernoulli negatives, margin 5, L1, normalised, SGD lr 1.0",
              "E": E, "R": R, "D": D, "triples_per_rank_per_step": n, "parallelism": f"dp{world}", "l2": "256 MiB buffer written between timed steps (L2 flush)"}
    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        with stdout_to_stderr():
            cpu_reference_step(w, threads, 1)
            s, kind = cpu_reference_step(w, threads, max(1, min(args.steps, 5)))
        v = n / s
        args.emit({"impl": "reference", "metric": "train triples/sec", "value": v, "unit": "triples/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * s, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "triples/s", "cores": threads, "kind": kind,
                                           "sample": f"{min(args.steps, 5)} full steps: Base.so sampling + torch-CPU forward/backward"},
                          "e2e": {"value": v, "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    import torch
    import torch.distributed as dist
    import mre_b200
    mre_b200.build()
    eng = mre_b200.engine
    torch.cuda.set_device(local)
    dctx = mre_b200.dist.DistContext() if world > 1 else None
    dev = torch.device("cuda", local)
    ctx = eng.Context(local)
    ix = eng.KGIndex.from_arrays(E, R, *splits).to_device(local)
    smp = eng.Sampler(ix, ctx=ctx, seed=192, stream_id=rank)
    # both tables (and both gradient tables) live back to back in ONE buffer.  N > 1: the buffers sit in NVLink peer memory and
    # the gradient sum is FUSED into the SGD kernel (mre_dp_sgd_step: reduce-scatter + update + all-gather over P2P loads / stores,
    # 1/world folded into the learning rate); --opt train_allreduce=nccl (or a box without CUDA IPC) falls back to one NCCL
    # all-reduce of the flat buffer + the plain SGD kernel, and the line says which
    n_par = (E + R) * D
    pg, exchange = None, "none (1 GPU)"
    want_nccl = any(o == "train_allreduce=nccl" for o in getattr(args, "opt", []) or [])
    if world > 1 and not want_nccl:
        ok = torch.ones(1, device=dev)
        try:
            pg = mre_b200.dist.PeerGroup(ctx, n_par)
        except Exception as e:  # noqa: BLE001
            sys.stderr.write(f"[rank {rank}] peer memory unavailable ({e}); falling back to NCCL\n")
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            pg = None
    if pg is not None:
        tab, grad = pg.weights, pg.grads
        exchange = "fused: gradient reduce-scatter + SGD + weight all-gather in ONE kernel over NVLink peer memory (mre_dp_sgd_step)"
    else:
        tab = torch.empty(n_par, dtype=torch.float32, device=dev)
        grad = torch.zeros_like(tab)
        if world > 1:
            exchange = "NCCL all-reduce of the flat gradient buffer + SGD kernel"
    ent_d, rel_d = tab[:E * D].view(E, D), tab[E * D:].view(R, D)
    g_ent, g_rel = grad[:E * D].view(E, D), grad[E * D:].view(R, D)
    ent_d.copy_(torch.from_numpy(ent)); rel_d.copy_(torch.from_numpy(rel))
    state = {"step": 0}
    lr = 1.0

    # The next batch does not depend on the weights: its Philox sampling runs on a side stream BESIDE this step's gradient
    # exchange / update kernel (which leaves most of every SM free) and is joined by an event before the next forward.
    side = torch.cuda.Stream(device=dev)
    bufs = [tuple([torch.empty(n, dtype=torch.int64, device=dev) for _ in range(3)] + [torch.empty(n, dtype=torch.float32, device=dev)])
            for _ in range(2)]                       # two batches in flight: the one being trained on and the one being drawn
    free_ev = [None, None]                           # recorded on the main stream when a buffer's batch has been consumed

    def draw():
        k = state["step"] & 1
        with torch.cuda.stream(side):
            if free_ev[k] is not None:
                side.wait_event(free_ev[k])
            batch = smp.sample(state["step"], B, neg, out=bufs[k])
            ev = torch.cuda.Event()
            ev.record(side)
        state["step"] += 1
        return batch, ev, k

    def step_dev():
        if state.get("next") is None:
            state["next"] = draw()
        (h, t, r, y), ev, k = state["next"]
        torch.cuda.current_stream().wait_event(ev)
        loss, _, _, _ = eng.transe_margin_step(ctx, ent_d, rel_d, h, t, r, B, neg, 5.0, 1, True, grad_ent=g_ent, grad_rel=g_rel)
        free_ev[k] = torch.cuda.Event()
        free_ev[k].record()
        state["next"] = draw()
        update()
        return loss

    def update():
        if pg is not None:
            pg.sgd_step(lr / world)
            return
        if dctx is not None:
            dctx.all_reduce_flat(grad)
        eng.sgd_update(ctx, tab, grad, lr / world)

    host = [np.empty(n, np.int64) for _ in range(3)] + [np.empty(n, np.float32)]
    pinned = [torch.empty(n, dtype=torch.int64).pin_memory() for _ in range(3)]

    def step_e2e():
        # host batch in (the reference's loader contract), loss out
        smp.sample_host(state["step"], B, neg, out=tuple(host))
        state["step"] += 1
        for p, a in zip(pinned, host[:3]):
            p.copy_(torch.from_numpy(a))
        h, t, r = (p.to(dev, non_blocking=True) for p in pinned)
        loss, _, _, _ = eng.transe_margin_step(ctx, ent_d, rel_d, h, t, r, B, neg, 5.0, 1, True, grad_ent=g_ent, grad_rel=g_rel)
        update()
        return float(loss.item())

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()                     # every rank's buffers hold the initial tables before anyone's first exchange
    for _ in range(args.warmup):
        step_dev()
    barrier()
    ctx.timing(True)
    ctx.timing_read()
    l0 = ctx.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local) as clocks:
        for a, b in ev:
            flush.zero_()
            a.record()
            loss = step_dev()
            b.record()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        launches = ctx.launches - l0
        kern_ms, kern_n = ctx.timing_read()
        ctx.timing(False)
        for _ in range(args.warmup):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([ms, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = t.tolist()
    value = world * n * args.steps / (ms * 1e-3)
    kms = kern_ms / max(kern_n, 1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = 9.0 * 4 * D * n      # 3 rows read in forward, 3 in backward, 3 rows of gradient atomics, per triple
    ach = alg_bytes / (kms * 1e-3) / 1e9
    cpu_base = None
    if rank == 0 and not args.no_extra:
        threads = os.cpu_count() or 1
        with stdout_to_stderr():
            s, kind = cpu_reference_step(w, threads, 3)
        cpu_base = {"value": n / s, "unit": "triples/s", "cores": threads, "kind": kind,
                    "sample": "3 full steps: Base.so sampling on all host threads + torch-CPU forward/backward of the reference expressions"}
    xinfo = {"gradient_exchange": exchange}
    if pg is not None:
        pg.check()
        wsum = tab.double().sum().reshape(1)                     # the replicas must hold identical weights
        ws = [torch.zeros_like(wsum) for _ in range(world)]
        dist.all_gather(ws, wsum)
        xinfo["replicas_identical"] = bool(all(torch.equal(x, ws[0]) for x in ws))
    if rank == 0:
        args.emit({
            "metric": "train triples/sec", "value": value, "unit": "triples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None,
                         "kernel": "transe fwd + loss + grad + bwd kernels (one timed group per step)", "kernel_ms": kms,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                         "algorithmic": "9 rows x 4D bytes per triple (3 read fwd, 3 read bwd, 3 gradient rows of atomics); tables are L2-resident (11.8 MB)"},
            "cpu_baseline": cpu_base,
            "e2e": {"value": world * n * args.steps / e2e_s, "unit": "triples/s", "h2d_bytes_per_step": 3 * n * 8, "d2h_bytes_per_step": 3 * n * 8 + n * 4 + 4,
                    "api": "sample_host (reference loader contract: numpy batch on the host) -> pinned H2D -> margin step -> loss.item()"},
            "gpu_launches": int(launches), "clocks": clocks.summary(), "result": {"last_loss": float(loss.item())}, **xinfo})
    if world > 1 and not getattr(args, "nested", False):
        dist.destroy_process_group()
