#!/usr/bin/env python
"""Attribute ncu warp-state samples / executed instructions of a kernel to its (noinline) device functions.
    python tools/ncu_by_function.py prof.ncu-rep lib.so [kernel-substring]
Uses nvdisasm's function labels of the matching cubin (offsets inside .text.<kernel>) and ncu's SASS source page."""
import csv, io, os, re, subprocess, sys, tempfile
rep, lib = sys.argv[1], sys.argv[2]
want = sys.argv[3] if len(sys.argv) > 3 else "k_sweep"
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
funcs, cur, insec = [], None, False
for line in dis.splitlines():
    m = re.match(r"^(\$?[_A-Za-z][\w\$]*):\s*$", line)
    if m and not m.group(1).startswith(".L"):
        name = m.group(1)
        insec = want in name
        cur = name if insec else None
        continue
    mm = re.match(r"^\s+/\*([0-9a-f]{4,6})\*/", line)
    if mm and cur:
        off = int(mm.group(1), 16)
        if not funcs or funcs[-1][1] != cur:
            funcs.append((off, cur))
def short(n):
    s = n.split("$")[-1] if "$" in n else n
    return subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip().split("(")[0]
starts = [(o, short(n)) for o, n in funcs]
src = subprocess.run(["ncu", "-i", rep, "--kernel-name", "regex:" + want, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_")] 
data = []
for r in rows[h + 1:]:
    if len(r) > iex and r[ia].startswith("0x"):
        data.append((int(r[ia], 16), int(r[isamp] or 0), int(r[iex] or 0), r))
    elif r and r[0] == "Kernel Name":
        break
base = data[0][0]
agg = {}
for a, sm, ex, r in data:
    off = a - base
    name = "?"
    for o, n in starts:
        if o <= off: name = n
        else: break
    t = agg.setdefault(name, [0, 0, {}])
    t[0] += sm; t[1] += ex
    for i in stall_cols:
        try: v = int(r[i] or 0)
        except ValueError: v = 0
        if v: t[2][hdr[i]] = t[2].get(hdr[i], 0) + v
ts, te = sum(v[0] for v in agg.values()) or 1, sum(v[1] for v in agg.values()) or 1
print("%-28s %8s %6s %12s %6s  top stalls" % ("function", "samples", "%", "inst_exec", "%"))
for n, (sm, ex, st) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    top = ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100.0 * v / max(sm, 1)) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print("%-28s %8d %5.1f%% %12d %5.1f%%  %s" % (n[:28], sm, 100.0 * sm / ts, ex, 100.0 * ex / te, top))
