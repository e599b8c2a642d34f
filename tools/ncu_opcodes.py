#!/usr/bin/env python
"""Executed warp-instructions of a kernel by SASS opcode, plus the hottest SASS lines (ncu source page, read here).
    python tools/ncu_opcodes.py prof.ncu-rep [nlines]"""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]; nl = int(sys.argv[2]) if len(sys.argv) > 2 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
c, s, tot, ts, lines = Counter(), Counter(), 0, 0, []
for r in rows[h + 1:]:
    if not (len(r) > iex and r[0].startswith("0x")):
        if r and r[0] == "Kernel Name": break
        continue
    n, sm = int(r[iex] or 0), int(r[ismp] or 0)
    t = r[isrc].split()
    op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")).split(".")[0]
    c[op] += n; s[op] += sm; tot += n; ts += sm
    lines.append((sm, n, r[0], r[isrc]))
print("total warp instructions %d, samples %d" % (tot, ts))
for k, v in c.most_common(22):
    print("%-10s %12d %5.1f%%   samples %5.1f%%" % (k, v, 100.0 * v / tot, 100.0 * s[k] / max(ts, 1)))
if nl:
    print("hottest lines by samples:")
    for sm, n, a, t in sorted(lines, reverse=True)[:nl]:
        print("%7d %10d %s %s" % (sm, n, a, t[:90]))
