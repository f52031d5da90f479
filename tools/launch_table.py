#!/usr/bin/env python3
"""Per-kernel device time from an `ncu --metrics gpu__time_duration.sum --csv` log.
usage: launch_table.py launches.csv [tail N launches | 0 = all] [--list]"""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
data = []
for r in rows[start:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    data.append((r[ki], v))
tail = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if tail:
    data = data[-tail:]
if "--list" in sys.argv:
    for n, v in data:
        print(f"{v:9.1f} us  {n[:110]}")
tot, cnt = defaultdict(float), defaultdict(int)
for n, v in data:
    k = n.split("(")[0][:80]
    tot[k] += v
    cnt[k] += 1
T = sum(tot.values())
print(f"total {T:.1f} us over {len(data)} launches")
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:25]:
    print(f"{v:10.1f} us {100 * v / T:5.1f}%  n={cnt[k]:4d}  {k}")
