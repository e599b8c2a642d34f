import os, sys, ctypes as C
sys.path.insert(0, "/root/repo")
from w2dynamics_b200 import problems
from w2dynamics_b200.CTQMC import MCTQMC
for same in (0, 1):
    p = problems.make_problem("C4")
    eng = MCTQMC()
    eng.set_option("nchains", 32768); eng.set_option("nbins", 32); eng.set_option("debug_same_key", same)
    eng.init_rnd(43890); eng.init_paras(p.pstring()); eng.init_solver(p.u_matrix, p.ftau, p.muimp, None, 1)
    eng.set_taudiffmax(p.qmc["TaudiffMax"]); eng.ctqmc_warmup(3000); eng.init_counters()
    L = eng._load(); ms = (C.c_double * 3)()
    eng._ck(L.w2b_bench_blocks(C.c_longlong(3), C.c_longlong(100), C.c_int(1), ms))
    eng.set_option("time_kernels", 1)
    eng._ck(L.w2b_bench_blocks(C.c_longlong(10), C.c_longlong(100), C.c_int(1), ms))
    print("same_key", same, "ms/step", ms[0] / 10, {k: eng.scalar("ms_" + k) / 10 for k in ("propose", "trace", "decide")}, "outer dim mix / order", eng.scalar("mean_order"))
    eng.dest_ctqmc()
