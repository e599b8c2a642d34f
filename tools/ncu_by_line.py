#!/usr/bin/env python
"""Attribute executed warp-instructions / stall samples of a kernel to source lines (nvdisasm -g line table + ncu SASS page).
    python tools/ncu_by_line.py prof.ncu-rep lib.so kernel-substring [ntop]"""
import csv, io, os, re, subprocess, sys, tempfile
rep, lib, want = sys.argv[1], sys.argv[2], sys.argv[3]
ntop = int(sys.argv[4]) if len(sys.argv) > 4 else 40
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-c", "-g", os.path.join(d, cubin)], capture_output=True, text=True).stdout
line_of, cur, insec, where = {}, None, False, ("?", 0)
for line in dis.splitlines():
    m = re.match(r"^\s*//## File \"([^\"]+)\", line (\d+)", line)
    if m:
        where = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"^\.text\.(\S+):", line)
    if m:
        insec = want in m.group(1); continue
    mm = re.match(r"^\s+/\*([0-9a-f]{4,6})\*/", line)
    if mm and insec:
        line_of[int(mm.group(1), 16)] = where
src = subprocess.run(["ncu", "-i", rep, "--kernel-name", "regex:" + want, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
iex, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [(int(r[0], 16), int(r[iex] or 0), int(r[ismp] or 0)) for r in rows[h + 1:] if len(r) > iex and r[0].startswith("0x")]
base = data[0][0]
agg = {}
for a, n, sm in data:
    w = line_of.get(a - base, ("?", 0))
    t = agg.setdefault(w, [0, 0]); t[0] += n; t[1] += sm
tot, ts = sum(v[0] for v in agg.values()) or 1, sum(v[1] for v in agg.values()) or 1
print("total instr %d samples %d" % (tot, ts))
cache = {}
def text(f, l):
    for root in ("w2dynamics_b200/csrc", "."):
        p = os.path.join(root, f)
        if os.path.exists(p):
            if p not in cache: cache[p] = open(p).read().splitlines()
            return cache[p][l - 1].strip()[:110] if 0 < l <= len(cache[p]) else ""
    return ""
for (f, l), (n, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:ntop]:
    print("%10d %5.1f%% smp %5.1f%%  %s:%d  %s" % (n, 100.0 * n / tot, 100.0 * sm / ts, f, l, text(f, l)))
