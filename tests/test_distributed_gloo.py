"""Multi-process path on CPU (gloo, world_size 2): chains are sharded over ranks by global chain id and the raw
accumulators are combined with ONE all-reduce per ctqmc_measure; the result must equal a
single process owning all chains (same Philox keys -> identical Markov chains -> bit-identical integer sums).
Uses the host-emulation build of the kernel logic; on the GPU box the same code path runs over NCCL."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch.distributed as dist
import _hostemu
from w2dynamics_b200 import problems
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
p = problems.make_problem({cfg!r}, nftau=1001, Nmeas=20, Ntau=50, NLegMax=10, MeasGiw=1, Niw=12)
eng = _hostemu.make()
eng.set_option("nchains", 4); eng.set_option("nbins", 4)
eng.enable_distributed()
eng.init_rnd(43890); eng.init_paras(p.pstring()); eng.init_solver(p.u_matrix, p.ftau, p.muimp, None, 1)
eng.set_taudiffmax(p.qmc["TaudiffMax"]); eng.ctqmc_warmup(500); eng.ctqmc_measure(1, 1)
np.savez({out!r} % dist.get_rank(), gtau=eng.gtau, occ=eng.occ, histo=eng.histo, sign=eng.mean_sign, gleg=eng.gleg, dm=eng.densitymatrix, giw=eng.giw)
dist.barrier(); dist.destroy_process_group()
"""


@pytest.mark.parametrize("cfg", ["C1", "C3"])            # 1x1 blocks (CTA scheduler) and blocked traces (round kernels, density matrix)
def test_two_rank_allreduce_equals_single_process(built, tmp_path, cfg):
    import _hostemu
    from w2dynamics_b200 import problems
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "r%d.npz")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port, out=out, cfg=cfg))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    for pr in procs:
        o, _ = pr.communicate(timeout=600)
        assert pr.returncode == 0, o.decode()[-2000:]
    r0, r1 = np.load(out % 0), np.load(out % 1)
    for k in ("gtau", "occ", "histo", "sign", "gleg", "dm", "giw"):
        assert np.array_equal(r0[k], r1[k]), k                      # every rank holds the reduced result
    # single process with all 8 chains (global chain ids 0..7)
    p = problems.make_problem(cfg, nftau=1001, Nmeas=20, Ntau=50, NLegMax=10, MeasGiw=1, Niw=12)
    eng = _hostemu.make()
    eng.set_option("nchains", 8); eng.set_option("nbins", 4)
    eng.init_rnd(43890); eng.init_paras(p.pstring()); eng.init_solver(p.u_matrix, p.ftau, p.muimp, None, 1)
    eng.set_taudiffmax(p.qmc["TaudiffMax"]); eng.ctqmc_warmup(500); eng.ctqmc_measure(1, 1)
    assert np.array_equal(eng.histo, r0["histo"]) and eng.mean_sign == float(r0["sign"])
    assert np.abs(eng.gtau - r0["gtau"]).max() < 1e-12 * max(1.0, np.abs(eng.gtau).max())
    assert np.abs(eng.occ - r0["occ"]).max() < 1e-12
    assert np.abs(eng.densitymatrix - r0["dm"]).max() < 1e-12
    assert r0["giw"].shape == (p.nbands, 2, 24) and np.abs(eng.giw - r0["giw"]).max() < 1e-12 * np.abs(eng.giw).max()   # the Matsubara sums ride in the same slab
    eng.dest_ctqmc()
