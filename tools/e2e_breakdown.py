#!/usr/bin/env python
"""Where an end-to-end plugin solve of ONE accumulation block spends its time (bench.py's e2e step), call by call.
    python tools/e2e_breakdown.py [C4] [nchains]"""
import os, sys, time
from collections import defaultdict
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from w2dynamics_b200 import problems
from w2dynamics_b200.CTQMC import MCTQMC
from w2dynamics_b200.solver_sequence import PluginSequence, qmc_defaults

cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
p = problems.make_problem(cfg)
eng = MCTQMC()
eng.set_option("nchains", nc); eng.set_option("nbins", 32); eng.set_option("device", 0); eng.set_option("reuse_population", 1)
times = defaultdict(float)


class Timed:
    def __init__(self, t): object.__setattr__(self, "_t", t)
    def __getattr__(self, name):
        t0 = time.perf_counter(); v = getattr(self._t, name); dt = time.perf_counter() - t0
        if callable(v):
            def call(*a, **k):
                t0 = time.perf_counter(); r = v(*a, **k); self._t.sync() if name in ("init_solver", "ctqmc_warmup", "set_taudiffmax") else None
                times[name] += time.perf_counter() - t0; return r
            return call
        times["read " + name] += dt
        return v
    def __setattr__(self, name, value): setattr(self._t, name, value)


qmc = qmc_defaults(); qmc.update(p.qmc)
ftau5 = np.ascontiguousarray(p.ftau.transpose(4, 0, 1, 2, 3)); scr = np.zeros_like(ftau5)
def solve(nwarm, nmeas, it, timed):
    q = dict(qmc); q.update(Nwarmups=nwarm, Nmeas=nmeas)
    seq = PluginSequence(Timed(eng) if timed else eng, q, 43890)
    seq.set_problem(p.hamiltonian, p.beta, p.nbands, p.quantum_numbers, p.u_matrix, ftau5, -p.muimp, scr)
    return seq.solve(it)
solve(3000, 1, 0, False)
for it in range(3): solve(0, 1, 1 + it, False)
times.clear(); n = 10
t0 = time.perf_counter()
for it in range(n): solve(0, 1, 10 + it, True)
tot = (time.perf_counter() - t0) / n
print("%s: %.1f ms per solve" % (cfg, tot * 1e3))
for k, v in sorted(times.items(), key=lambda kv: -kv[1])[:14]:
    print("  %-28s %7.2f ms" % (k, v / n * 1e3))
