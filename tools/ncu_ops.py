"""Opcode histogram of one kernel from an `ncu --page source --csv` dump (development aid)."""
import collections
import csv
import re
import sys

path, per = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
hdr = next(r for r in rows if r and r[0] == "Address")
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
si, ie, ss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[ie]) for r in data)
ops, samp = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si])
    op = m.group(2).split(".")[0] if m else "?"
    ops[op] += int(r[ie])
    samp[op] += int(r[ss])
ts = sum(samp.values())
print(f"total warp instructions {tot}, per unit {tot / per:.1f}")
for op, c in ops.most_common(30):
    print(f"{op:10s} {c / tot:6.1%}  per unit {c / per:8.1f}  samples {samp[op] / ts:6.1%}")
