"""Developer helper: one line per bench JSON + per-kernel averages of ncu launch-list CSVs (gpu__time_duration.sum)."""
import collections
import csv
import glob
import json
import sys


def bench_lines(pattern):
    for f in sorted(glob.glob(pattern)):
        try:
            j = json.loads(open(f).read())
        except Exception:  # noqa: BLE001
            print(f, "ERR", open(f.replace(".json", ".err")).read()[-600:])
            continue
        r = j["roofline"]
        print(f"{f.split('/')[-1]:28s} value={j['value']:.3e} ms/step={j['ms_per_step']:.4f} kernel_ms={r['kernel_ms']:.4f} "
              f"step/kernel={j['ms_per_step'] / r['kernel_ms']:.3f} frac={r['frac']:.3f} pipe={r.get('pipe_frac')} "
              f"resc={r.get('rescored_fraction')} e2e={j['e2e']['value']:.3e} launches={j['gpu_launches']}")


def launch_list(fn):
    print(fn)
    hdr, agg = None, collections.OrderedDict()
    for r in csv.reader(open(fn)):
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            agg.setdefault(d["Kernel Name"][:70], []).append(v)
    for k, v in agg.items():
        print(f"  {k:70s} n={len(v):3d} avg={sum(v) / len(v) / 1000:8.2f} us")


if __name__ == "__main__":
    for a in sys.argv[1:]:
        (launch_list if a.endswith(".csv") else bench_lines)(a)
