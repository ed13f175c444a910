"""Developer timing probe: host index build (mre_index_create + mre_index_to_device) vs the GPU build (mre_index_create_device)."""
import json, os, sys, time
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import numpy as np
import mre_b200
eng = mre_b200.engine
out = []
SIZES = ((310_116, 14_541, 237), (5_200_000, 2_000_000, 1_000), (50_000_000, 20_000_000, 10_000))
for n, E, R in SIZES[int(os.environ.get("MRE_IB_FIRST", "0")):]:
    rng = np.random.default_rng(1)
    tr = tuple(rng.integers(0, m, n) for m in (E, E, R))
    va = tuple(rng.integers(0, m, n // 20) for m in (E, E, R))
    te = tuple(rng.integers(0, m, n // 20) for m in (E, E, R))
    t0 = time.perf_counter(); dev = eng.KGIndex.from_arrays_device(E, R, tr, va, te, device=0); t1 = time.perf_counter()
    dev2 = eng.KGIndex.from_arrays_device(E, R, tr, va, te, device=0); t2 = time.perf_counter()
    host = eng.KGIndex.from_arrays(E, R, tr, va, te); t3 = time.perf_counter()
    host.to_device(0); t4 = time.perf_counter()
    same = all(np.array_equal(host.device_column(c), dev2.device_column(c), equal_nan=True) for c in range(11))
    bits = lambda m: max(1, int(np.ceil(np.log2(m))))
    per_sort = 2 * -(-bits(E) // 8) + -(-bits(R) // 8)
    n_all = n + 2 * (n // 20)
    # six sorts: all (n_all) x 2 orders, train (n) x 2 orders, test, valid; one pass reads the key twice and moves three int32 columns
    elems = per_sort * (2 * n_all + 2 * n + 2 * (n // 20))
    out.append({"triples": n_all, "E": E, "R": R, "passes_per_sort": per_sort, "gpu_build_ms_device": dev2.build_ms,
                "gpu_build_s_wall_first": t1 - t0, "gpu_build_s_wall": t2 - t1, "host_build_s": t3 - t2, "host_upload_s": t4 - t3,
                "sort_traffic_gb": elems * 32 / 1e9, "same_bits": bool(same)})
    print(json.dumps(out[-1]), flush=True)
    del dev, dev2, host
