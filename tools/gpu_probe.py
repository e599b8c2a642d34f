#!/usr/bin/env python
"""Quick GPU probe: throughput of the sweep / measure kernels for a config, chain counts and kernel geometries.
    python tools/gpu_probe.py C2 16384 3000 10 [cta_warps,batch_chains ...]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from w2dynamics_b200 import problems
from w2dynamics_b200.CTQMC import MCTQMC

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
chains = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["1024", "4096", "16384"])]
nwarm = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
nblk = int(sys.argv[4]) if len(sys.argv) > 4 else 20
geoms = [tuple(int(v) for v in g.split(",")) for g in sys.argv[5:]] or [(0, 0)]
p = problems.make_problem(cfg)
for nc in chains:
    for (cw, bc) in geoms:
        eng = MCTQMC()
        eng.set_option("nchains", nc); eng.set_option("nbins", 32)
        if cw: eng.set_option("cta_warps", cw)
        if bc: eng.set_option("batch_chains", bc)
        if os.environ.get("W2B_SAME_KEY"): eng.set_option("debug_same_key", 1)
        if os.environ.get("W2B_PRUN"): eng.set_option("propose_run", int(os.environ["W2B_PRUN"]))
        if os.environ.get("W2B_PREFETCH"): eng.set_option("prefetch", int(os.environ["W2B_PREFETCH"]))
        eng.init_rnd(43890); eng.init_paras(p.pstring())
        t0 = time.time()
        eng.init_solver(p.u_matrix, p.ftau, p.muimp, None, 1)
        eng.set_taudiffmax(p.qmc["TaudiffMax"])
        t1 = time.time()
        eng.ctqmc_warmup(nwarm)
        t2 = time.time()
        eng.init_counters()
        eng.sync(); t3 = time.time()
        for _ in range(nblk):
            eng.block(100, False)
        eng.sync(); t4 = time.time()
        for _ in range(nblk):
            eng.block(100, True)
        eng.sync(); t5 = time.time()
        eng.block(1000, False); eng.sync(); t6 = time.time()
        print("cfg %s chains %6d W %2d ncb %3d grid %4d smem %6d: warmup(%d) %.2fs | sweep100 %.3e | sweep100+meas %.3e | sweep1000 %.3e moves/s | max_k %d err %d"
              % (cfg, nc, eng.scalar("cta_warps"), eng.scalar("batch_chains"), eng.scalar("sweep_grid"), eng.scalar("sweep_smem"), nwarm, t2 - t1,
                 nc * 100 * nblk / (t4 - t3), nc * 100 * nblk / (t5 - t4), nc * 1000 / (t6 - t5),
                 eng.scalar("max_k"), eng.scalar("error_flags")), flush=True)
        eng.dest_ctqmc()
