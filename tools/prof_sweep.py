#!/usr/bin/env python
"""Profiling driver: thermalise, then a few 100-move sweep launches (+ optional measurement) for ncu to capture.
    ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 3 -c 1 -o gpurun_out/prof python tools/prof_sweep.py C4 8192"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from w2dynamics_b200 import problems
from w2dynamics_b200.CTQMC import MCTQMC

cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
nwarm = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
p = problems.make_problem(cfg)
eng = MCTQMC()
eng.set_option("nchains", nc); eng.set_option("nbins", 32)
for k in ("cta_warps", "batch_chains", "propose_run", "prefetch", "lu_groups", "scheduler"):
    if os.environ.get("W2B_" + k.upper()):
        eng.set_option(k, int(os.environ["W2B_" + k.upper()]))
eng.init_rnd(43890); eng.init_paras(p.pstring())
eng.init_solver(p.u_matrix, p.ftau, p.muimp, None, 1)
eng.set_taudiffmax(p.qmc["TaudiffMax"])
eng.ctqmc_warmup(nwarm)
eng.init_counters()
for _ in range(6):
    eng.block(100, True)
eng.sync()
print("done", cfg, nc, "W", eng.scalar("cta_warps"), "ncb", eng.scalar("batch_chains"), "grid", eng.scalar("sweep_grid"), "smem", eng.scalar("sweep_smem"))
eng.dest_ctqmc()
