#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) per kernel: launches, total, mean, share.

    python tools/launch_summary.py gpurun_out/r02c_launches.csv > profiles/r02c_launches_summary.md
"""
import csv, sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
col = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0.0])
for r in rd:
    if len(r) < len(hdr) or r[col["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[col["Kernel Name"]]
    val = float(r[col["Metric Value"]].replace(",", ""))
    unit = r[col["Metric Unit"]]
    us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
    short = name.split("(")[0].strip()
    if short.startswith("void "):
        short = short[5:]
    agg[short][0] += 1
    agg[short][1] += us
ours = {k: v for k, v in agg.items() if k.startswith("lcr::")}
tot = sum(v[1] for v in ours.values())
print("| kernel | launches | total ms | mean us | share of liblcr time |")
print("|---|---|---|---|---|")
for k, (n, us) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {us / 1e3:.3f} | {us / n:.1f} | {100 * us / tot:.1f} % |")
other = sum(v[1] for k, v in agg.items() if not k.startswith("lcr::"))
print(f"\nliblcr total {tot / 1e3:.2f} ms over {sum(v[0] for v in ours.values())} launches; other kernels (ATen input generation, copies) {other / 1e3:.2f} ms over "
      f"{sum(v[0] for k, v in agg.items() if not k.startswith('lcr::'))} launches.")
