#!/usr/bin/env python
"""Small runs of every move type and scheduler for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from w2dynamics_b200 import problems
from w2dynamics_b200.CTQMC import MCTQMC

CASES = [  # cfg, beta, scheduler, warm-up moves, extra QMC keys
    ("C2", 4.0, 0, 3000, dict(Percentage4OperatorMove=0.2, PercentageOuterMove=0.2, PercentageGlobalMove=0.05, PercentageTauShiftMove=0.05)),
    ("C2", 4.0, 2, 2000, dict(Percentage4OperatorMove=0.2, PercentageOuterMove=0.2, PercentageGlobalMove=0.05, PercentageTauShiftMove=0.05)),
    ("C2", 4.0, 3, 2000, dict(Percentage4OperatorMove=0.2, PercentageOuterMove=0.2, PercentageGlobalMove=0.05, PercentageTauShiftMove=0.05)),
    ("C4", 2.0, 0, 1500, dict(Percentage4OperatorMove=0.2, PercentageOuterMove=0.2, PercentageGlobalMove=0.05, PercentageTauShiftMove=0.05)),
    ("C4", None, 0, 1500, dict()),
    ("C5", None, 0, 5000, dict(PercentageGlobalMove=0.02)),          # orders >= 24: blocked Gauss-Jordan
    ("C5", None, 3, 5000, dict(PercentageGlobalMove=0.02, Percentage4OperatorMove=0.02)),
]
sel = os.environ.get("SANITIZE_CASES")
if sel:
    CASES = [CASES[int(i)] for i in sel.split(",")]
scale = float(os.environ.get("SANITIZE_SCALE", "1"))
for cfg, beta, sched, warm, extra in CASES:
    warm = int(warm * scale)
    p = problems.make_problem(cfg, nftau=1001, beta=beta, Nmeas=3, **extra)
    eng = MCTQMC()
    eng.set_option("nchains", 48); eng.set_option("nbins", 4); eng.set_option("scheduler", sched); eng.set_option("reuse_population", 1)
    eng.init_rnd(43890); eng.init_paras(p.pstring()); eng.init_solver(p.u_matrix, p.ftau, p.muimp, None, 1)
    eng.set_taudiffmax(p.qmc["TaudiffMax"])
    eng.ctqmc_warmup(warm)
    eng.ctqmc_measure(1, 1)
    k = eng.scalar("mean_order")
    eng.dest_ctqmc()
    eng.init_paras(p.pstring()); eng.init_solver(p.u_matrix, p.ftau, p.muimp, None, 2)      # revive: k_rebuild
    eng.set_taudiffmax(p.qmc["TaudiffMax"]); eng.ctqmc_warmup(200); eng.ctqmc_measure(1, 1)
    print(cfg, beta, "scheduler", sched, "mean order %.1f" % k, "acc4 %.3f" % float(eng.accadd4) if extra.get("Percentage4OperatorMove") else "", flush=True)
    eng.set_option("reuse_population", 0); eng.dest_ctqmc(); eng.dest_paras()
print("done")
