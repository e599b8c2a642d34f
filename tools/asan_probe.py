"""AddressSanitizer run of the kernel sources (host-emulation build, one lane per chain) over the opt-in features: state sampling,
outer pair moves, 4-operator moves with the rank-2 update, the Matsubara estimator; blocked and 1 x 1 configurations, phase and
autonomous schedulers.  Complements compute-sanitizer on the GPU (tools/sanitize_probe.py): same index arithmetic, host bounds.

    g++ -O1 -g -std=c++17 -x c++ -DW2B_HOST_EMU -fsanitize=address -fno-omit-frame-pointer -shared -fPIC \
        -o /tmp/asan/libw2b_hostemu_asan.so w2dynamics_b200/csrc/capi.cu
    LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tools/asan_probe.py /tmp/asan/libw2b_hostemu_asan.so
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from w2dynamics_b200 import problems
from w2dynamics_b200.CTQMC import MCTQMC
from test_boundary import _start
from test_state_sampling import lift_degeneracies

ALL = dict(statesampling=1, PercentageOuterMove=0.25, Percentage4OperatorMove=0.2)
for cfg, beta, extra, opts in [("C3", 3.0, dict(ALL, MeasGiw=1, Niw=13), {}), ("C4", 2.0, ALL, {}),
                               ("C2", 4.0, dict(ALL, MeasGiw=1, Niw=7), {"scheduler": 0}), ("C2", 4.0, ALL, {"scheduler": 3}),
                               ("C5", None, dict(Percentage4OperatorMove=0.3, MeasGiw=1, Niw=5), {})]:
    p = problems.make_problem(cfg, nftau=2001, beta=beta, PercentageGlobalMove=0.05, PercentageTauShiftMove=0.05, Nmeas=3, NCorr=50, **extra)
    if cfg in ("C3", "C4"):
        lift_degeneracies(p)
    eng = _start(MCTQMC(libpath=sys.argv[1]), p, nchains=6, nbins=2, **opts)
    eng.init_counters()
    eng.ctqmc_warmup(1500 if cfg != "C5" else 400)
    eng.ctqmc_measure(1, 1)
    print(cfg, opts, "error_flags", eng.scalar("error_flags"), "rank-2 updates", eng.scalar("cnt_four_rank2"), "finite", bool(np.isfinite(eng.gtau).all()))
    eng.dest_ctqmc()
print("no AddressSanitizer report")
