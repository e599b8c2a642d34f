#!/usr/bin/env python
"""Group a kernel's SASS by execution count (= loop nest regions) and print opcode mix; dumps the annotated SASS listing.
    python tools/ncu_regions.py prof.ncu-rep kernel-regex [out.txt]"""
import csv, io, subprocess, sys
from collections import Counter
rep, want = sys.argv[1], sys.argv[2]
src = subprocess.run(["ncu", "-i", rep, "--kernel-name", "regex:" + want, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
isrc, iex, ismp, ipr = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Predicated-On Thread Instructions Executed")
data, seen = [], set()
for r in rows[h + 1:]:
    if not (len(r) > iex and r[0].startswith("0x")):
        continue
    if r[0] in seen:
        break
    seen.add(r[0]); data.append((r[0], int(r[iex] or 0), int(r[ismp] or 0), r[isrc], int(r[ipr] or 0)))
tot = sum(d[1] for d in data); ts = sum(d[2] for d in data)
print("instructions %d (static %d), samples %d" % (tot, len(data), ts))
ops = Counter()
for a, n, s, t, p in data:
    w = t.split()
    op = (w[1] if w and w[0].startswith("@") and len(w) > 1 else (w[0] if w else "?")).split(".")[0]
    ops[op] += n
print("opcodes: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in ops.most_common(16)))
dm = [(n, p) for a, n, s, t, p in data if "DMMA" in t]
print("DMMA executed %d, predicated-on warp-equivalents %d -> %.1f instructions per DMMA" % (sum(n for n, p in dm), sum(p for n, p in dm) // 32, tot / max(1, sum(p for n, p in dm) // 32)))
reg = Counter(); smp = Counter()
for a, n, s, t, p in data:
    reg[n] += 1; smp[n] += s
for n, cnt in sorted(reg.items(), key=lambda kv: -kv[0] * kv[1])[:14]:
    print("count %10d x %4d instrs = %5.1f%% of instructions, %5.1f%% of samples" % (n, cnt, 100.0 * n * cnt / tot, 100.0 * smp[n] / max(ts, 1)))
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write("\n".join("%s %10d %6d %s" % d[:4] for d in data))
