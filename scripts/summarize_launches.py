#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, share.
usage: python scripts/summarize_launches.py gpurun_out/r01_launches.csv > profiles/r01_launch_shares.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, ui, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"<.*", "", name) if not name.startswith("void acan") and "acan" not in name else re.sub(r"\(anonymous namespace\)::|unnamed>::", "", r[ki])[:110]
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
ours = sum(v for k, v in tot.items() if "acan" in k or "gemm_kernel" in k or "head_kernel" in k or "decode_ord" in k)
print(f"# {len(rows)-1} launches, {total/1e3:.2f} ms of kernel time (cold-cache, serialised under ncu: compare SHARES, not absolutes)")
print(f"# hand-written acan kernels: {ours/1e3:.2f} ms = {100*ours/total:.1f} % of kernel time")
print(f"{'share%':>7} {'total_us':>10} {'n':>5} {'avg_us':>9}  kernel")
for k in sorted(tot, key=lambda k: -tot[k])[:45]:
    print(f"{100*tot[k]/total:7.2f} {tot[k]:10.1f} {cnt[k]:5d} {tot[k]/cnt[k]:9.1f}  {k}")
