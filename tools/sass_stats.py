"""Static SASS opcode histogram of one kernel in libhydpy_b200.so (development aid): python tools/sass_stats.py zone_kernel"""
import collections
import re
import subprocess
import sys

lib = "hydpy_b200/libhydpy_b200.so"
pat = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", out)
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if pat not in name:
        continue
    ops = collections.Counter()
    for line in b.split("\n"):
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(2).split(".")[0]] += 1
    total = sum(ops.values())
    fp64 = sum(ops[o] for o in ("DFMA", "DADD", "DMUL", "DSETP"))
    print(f"{name}: {total} instructions, FP64 {fp64}")
    print("  " + "  ".join(f"{o}:{c}" for o, c in ops.most_common(24)))
