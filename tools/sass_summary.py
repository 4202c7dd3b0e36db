#!/usr/bin/env python
"""Development: per-kernel SASS instruction mix of one object file (no GPU needed):

    python tools/sass_summary.py gmlm_b200/build/graphnorm.o [name-filter]

Prints, for every kernel whose demangled name contains the filter, the instruction count and the counts
of the mnemonics that matter for the roofline argument (128-bit global accesses, MUFU, FP64, shuffles)."""
import collections
import re
import subprocess
import sys

obj = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
KEYS = ["LDG.E.128", "LDG.E.64", "LDG.E", "STG.E.128", "STG.E", "LDS", "STS", "MUFU.RCP", "MUFU.EX2", "MUFU.RSQ",
        "DADD", "DFMA", "F2F.F64.F32", "FFMA", "FADD", "FMUL", "SHFL", "PRMT", "IMAD", "BAR.SYNC"]
cur, counts = None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"gmlm::\(anonymous namespace\)::", "", cur).split("(")[0].replace("void ", "")
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
                break
for name, c in counts.items():
    if flt in name:
        mix = "  ".join(f"{k}={c[k]}" for k in KEYS if c[k])
        print(f"{name[:100]}\n    {c['_total']} instructions   {mix}")
