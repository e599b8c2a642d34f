#!/usr/bin/env python
"""Static SASS instruction counts per source line for the functions whose name contains a substring (code-footprint diet).
    python tools/sass_by_line.py lib.so function-substring [ntop]"""
import os, re, subprocess, sys, tempfile
from collections import Counter
lib, want = sys.argv[1], sys.argv[2]
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 30
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-c", "-g", os.path.join(d, cubin)], capture_output=True, text=True).stdout
cnt, where, insec, tot = Counter(), ("?", 0), False, 0
for line in dis.splitlines():
    m = re.match(r"^\s*//## File \"([^\"]+)\", line (\d+)", line)
    if m:
        where = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"^(\$?[_A-Za-z][\w\$]*):\s*$", line)
    if m and not m.group(1).startswith(".L"):
        insec = want in m.group(1); continue
    if insec and re.match(r"^\s+/\*[0-9a-f]{4,6}\*/", line):
        cnt[where] += 1; tot += 1
print("total", tot)
cache = {}
def text(f, l):
    for root in ("w2dynamics_b200/csrc", "."):
        p = os.path.join(root, f)
        if os.path.exists(p):
            if p not in cache: cache[p] = open(p).read().splitlines()
            return cache[p][l - 1].strip()[:120] if 0 < l <= len(cache[p]) else ""
    return ""
for (f, l), n in cnt.most_common(ntop):
    print("%5d  %s:%d  %s" % (n, f, l, text(f, l)))
