#!/usr/bin/env python
"""Development: registers / spills per kernel from an `nvcc -Xptxas=-v` log (stdin or file)."""
import re
import subprocess
import sys

t = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
pat = re.compile(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores.*?\n.*?Used (\d+) registers", re.S)
for m in pat.finditer(t):
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"gmlm::\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name).split("(")[0]
    print(f"{m.group(4):>4} regs  {m.group(3):>4} spill  {name[:110]}")
