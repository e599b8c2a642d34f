#!/usr/bin/env python
"""Per-function SASS instruction counts of the kernels in the CUDA library (needs cuobjdump + nvdisasm; no GPU).
    python tools/sass_sizes.py [lib.so] [kernel-substring]"""
import os, re, subprocess, sys, tempfile
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "w2dynamics_b200", "libw2b_cthyb.so")
want = sys.argv[2] if len(sys.argv) > 2 else "k_sweep"
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    out = subprocess.run(["nvdisasm", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
cur, counts, order = None, {}, []
for line in out.splitlines():
    m = re.match(r"^(\$?[_A-Za-z][\w\$]*):\s*$", line)
    if m and not m.group(1).startswith(".L"):
        cur = m.group(1)
        if cur not in counts:
            counts[cur] = 0; order.append(cur)
        continue
    if cur and re.match(r"^\s+/\*[0-9a-f]{4,6}\*/", line):
        counts[cur] += 1
tot = 0
for name in order:
    if want in name:
        short = name.split("$")[-1] if "$" in name else name
        short = subprocess.run(["c++filt", short], capture_output=True, text=True).stdout.strip().split("(")[0]
        print("%6d  %s" % (counts[name], short)); tot += counts[name]
print("%6d  total (%s)" % (tot, want))
