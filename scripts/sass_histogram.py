#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of build/libmre_b200.so (cuobjdump -sass): the mnemonics that prove which hardware path a
kernel uses (UTCHMMA / UTCBAR / LDTM = tcgen05 + TMEM, UTMALDG = TMA, FADD2 / FFMA2 = packed FP32, SYNCS = mbarrier,
ATOM / RED = atomics).  Usage: python scripts/sass_histogram.py > profiles/rN_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multimodal-relation-extrapolation_b200", "build", "libmre_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "FADD2", "FFMA2", "FMUL2", "FADD", "FFMA", "FMUL", "FSETP",
        "MUFU", "LDS", "STS", "LDG", "STG", "ATOMG", "ATOMS", "RED", "REDG", "SHFL", "BAR", "IMAD", "LOP3", "POPC"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    filt = subprocess.run(["c++filt"], input=out, capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in filt.splitlines():
        m = re.match(r"\s*Function : (.*)", line)
        if m:
            cur = m.group(1).strip()
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
            kernels[cur]["_total"] += 1
    print("# SASS opcode histogram of libmre_b200.so (sm_100a), `cuobjdump -sass` -- instruction counts per kernel\n")
    print("Columns: every opcode of the key list that appears at least once in the kernel; `total` = all instructions.\n")
    for name, c in kernels.items():
        short = re.sub(r"\(.*", "", name)
        short = short.replace("mre::", "")
        cells = ", ".join(f"{k} {c[k]}" for k in KEYS if c[k])
        print(f"* `{short}` -- total {c['_total']}: {cells}")


if __name__ == "__main__":
    sys.exit(main())
