"""Developer helper: print the headline metrics + instruction mix + stall samples per code region of one .ncu-rep."""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_fma.avg.pct",
        "sm__pipe_fma_cycles_active.avg.pct", "sm__pipe_fmaheavy_cycles_active.avg.pct", "sm__inst_executed_pipe_alu.avg.pct",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "_per_issue_active.ratio", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor", "lts__t_bytes.sum",
        "lts__throughput.avg.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__grid_size", "launch__block_size"]
for h, u, v in zip(hdr, units, vals):
    if any(w in h for w in want):
        print(f"{h:95s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
tot = sum(int(r[ia]) for r in data); tots = sum(int(r[isamp]) for r in data)
c = Counter(); cs = Counter()
for r in data:
    t = r[isrc].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    c[op] += int(r[ia]); cs[op] += int(r[isamp])
print("total warp-instructions", tot, "samples", tots)
for op, n in c.most_common(14):
    print(f"  {op:12s} {100 * n / tot:6.2f}% of instructions   {100 * cs[op] / max(tots, 1):6.2f}% of samples")
seg, start, prev = [], 0, None
for i, r in enumerate(data):
    n = int(r[ia])
    if prev is None: prev = n
    if n != prev:
        seg.append((start, i - 1, prev)); start = i; prev = n
seg.append((start, len(data) - 1, prev))
for a, b, n in seg:
    s = sum(int(data[i][isamp]) for i in range(a, b + 1))
    if s > 0.004 * tots:
        print(f"  lines {a}-{b} executed {n}x: {100 * s / tots:5.2f}% of samples; first: {data[a][isrc][:60]}")
