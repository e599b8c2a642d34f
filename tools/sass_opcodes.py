#!/usr/bin/env python
"""Static SASS opcode histogram per kernel of the CUDA library (cuobjdump; no GPU): proves which hardware paths the binary uses
(DMMA = FP64 tensor pipe; LDL / STL = local-memory traffic; LDGSTS / UBLKCP / UTMALDG = async copies / TMA).
    python tools/sass_opcodes.py [lib.so] > profiles/rN_sass_opcode_counts.txt"""
import os, re, subprocess, sys
from collections import Counter, OrderedDict
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "w2dynamics_b200", "libw2b_cthyb.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, cur = OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cur = kern.setdefault(name, Counter()); continue
    m = re.match(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
watch = ["DMMA", "DFMA", "DMUL", "DADD", "LDG", "STG", "LDS", "STS", "LDL", "STL", "LDGSTS", "UBLKCP", "UTMALDG", "SHFL", "BAR", "ATOMG", "RED", "IMAD", "BRA"]
print("kernel".ljust(28) + "total".rjust(7) + "".join(w.rjust(8) for w in watch))
for name, c in kern.items():
    print(name[:27].ljust(28) + str(sum(c.values())).rjust(7) + "".join(str(c.get(w, 0)).rjust(8) for w in watch))
