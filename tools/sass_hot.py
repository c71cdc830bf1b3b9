#!/usr/bin/env python
"""Instruction mix of the hot part of a kernel from `ncu -i x.ncu-rep --page source --csv --kernel-name regex:<k>`.

  python tools/sass_hot.py <source.csv> <units> [min_executed]   units = divisor (e.g. warp-sub-steps) for the per-unit figures
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
units = float(sys.argv[2])
h = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[h]
ia, isrc, ismp, iex, ithr = (hdr.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed",
                                                      "Avg. Threads Executed"))
data = []
for r in rows[h + 1:]:
    if len(r) <= iex:
        continue
    try:
        data.append((r[ia], r[isrc].strip(), int(r[ismp]), int(r[iex]), r[ithr]))
    except ValueError:
        continue
tot = sum(d[3] for d in data)
thresh = float(sys.argv[3]) if len(sys.argv) > 3 else 0.2 * max(d[3] for d in data)
hot = [d for d in data if d[3] > thresh]
print(f"total executed {tot}, {len(data)} SASS rows; {len(hot)} hot rows hold {sum(d[3] for d in hot)} ({sum(d[3] for d in hot) / tot:.1%})")
mix = collections.Counter()
smp = collections.Counter()
for d in hot:
    parts = d[1].split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    mix[op.split(".")[0]] += d[3]
    smp[op.split(".")[0]] += d[2]
for k, v in mix.most_common():
    print(f"{k:10s} {v / units:8.2f} per unit   stall samples {smp[k]}")
print(f"{'all hot':10s} {sum(mix.values()) / units:8.2f} per unit")
if "--list" in sys.argv:
    for d in hot:
        print(f"{d[3]:12d} {d[2]:6d} {d[4]:>5s}  {d[1]}")
