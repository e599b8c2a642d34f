#!/usr/bin/env python
"""Where one warp's time goes in the autonomous sweep (diagnostics; needs a library built with -DW2B_TIMING):
    nvcc ... -DW2B_TIMING -o gpurun_out/libw2b_timing.so w2dynamics_b200/csrc/capi.cu ; python tools/tick_probe.py C5 32768 gpurun_out/libw2b_timing.so"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from w2dynamics_b200 import problems
from w2dynamics_b200.CTQMC import MCTQMC
cfg, nc, lib = sys.argv[1], int(sys.argv[2]), sys.argv[3]
p = problems.make_problem(cfg)
eng = MCTQMC(libpath=lib)
eng.set_option("nchains", nc); eng.set_option("nbins", 32); eng.set_option("scheduler", 3)
eng.init_rnd(43890); eng.init_paras(p.pstring()); eng.init_solver(p.u_matrix, p.ftau, p.muimp, None, 1)
eng.set_taudiffmax(p.qmc["TaudiffMax"])
eng.ctqmc_warmup(int(sys.argv[4]) if len(sys.argv) > 4 else 12000)
eng.init_counters()
eng.block(100, False); eng.sync()
names = ["add_propose", "add_decide", "rem_propose", "rem_decide", "global", "shift", "post_move"]
t = [eng.scalar("tick%d" % i) for i in range(7)]
cnt = {k: eng.scalar("cnt_" + k) for k in ("try_add", "try_rem", "acc_add", "acc_rem", "qn_add", "qn_rem", "try_glob", "try_shift")}
tot = sum(t)
print(cfg, nc, "mean order", eng.scalar("mean_order"), cnt)
for n, v in zip(names, t):
    print("  %-12s %6.2f %%   %.3e cycles" % (n, 100 * v / tot, v))
print("  cycles per move per warp: %.0f" % (tot / (nc * 100.0)))
