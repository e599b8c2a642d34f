"""Two ranks on two GPUs: chains shard over the ranks by global chain id and ctqmc_measure ends with ONE ncclAllReduce of the raw
accumulators inside the library (w2b_comm_init / w2b_allreduce_accum, include/w2b_cthyb.h) -- no torch collective on the data path.
The reduced result must equal a single process owning all chains (same Philox keys -> identical Markov chains -> the integer
sums agree bit for bit, the floating-point sums up to the order of the atomics).  Needs two visible GPUs; skipped otherwise
(the CPU twin of this test, tests/test_distributed_gloo.py, runs the same sharding logic over gloo)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
rank = int(sys.argv[1])
torch.cuda.set_device(rank)
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=2)     # only carries the 128-byte NCCL id
from w2dynamics_b200 import problems
from w2dynamics_b200.CTQMC import MCTQMC
p = problems.make_problem({cfg!r}, nftau=1001, Nmeas=20, Ntau=50, NLegMax=10)
eng = MCTQMC()
eng.set_option("nchains", 32); eng.set_option("nbins", 4); eng.set_option("device", rank)
eng.enable_distributed()
assert eng.scalar("comm_size") == 2 and eng.scalar("comm_rank") == rank
eng.init_rnd(43890); eng.init_paras(p.pstring()); eng.init_solver(p.u_matrix, p.ftau, p.muimp, None, 1)
eng.set_taudiffmax(p.qmc["TaudiffMax"]); eng.ctqmc_warmup(500); eng.ctqmc_measure(1, 1)
np.savez({out!r} % rank, gtau=eng.gtau, occ=eng.occ, histo=eng.histo, sign=eng.mean_sign, gleg=eng.gleg, dm=eng.densitymatrix)
eng.dest_ctqmc()
dist.barrier(); dist.destroy_process_group()
"""


@pytest.mark.parametrize("cfg", ["C2", "C4"])
def test_two_gpu_allreduce_equals_single_process(built, tmp_path, cfg):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from w2dynamics_b200 import problems
    from w2dynamics_b200.CTQMC import MCTQMC
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "r%d.npz")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port, out=out, cfg=cfg))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    for pr in procs:
        o, _ = pr.communicate(timeout=900)
        assert pr.returncode == 0, o.decode()[-3000:]
    r0, r1 = np.load(out % 0), np.load(out % 1)
    for k in ("gtau", "occ", "histo", "sign", "gleg", "dm"):
        assert np.array_equal(r0[k], r1[k]), k                      # every rank holds the reduced result
    p = problems.make_problem(cfg, nftau=1001, Nmeas=20, Ntau=50, NLegMax=10)
    eng = MCTQMC()
    eng.set_option("nchains", 64); eng.set_option("nbins", 4); eng.set_option("device", 0)
    eng.init_rnd(43890); eng.init_paras(p.pstring()); eng.init_solver(p.u_matrix, p.ftau, p.muimp, None, 1)
    eng.set_taudiffmax(p.qmc["TaudiffMax"]); eng.ctqmc_warmup(500); eng.ctqmc_measure(1, 1)
    assert np.array_equal(eng.histo, r0["histo"]) and abs(eng.mean_sign - float(r0["sign"])) < 1e-15
    assert np.abs(eng.gtau - r0["gtau"]).max() < 1e-10 * max(1.0, np.abs(eng.gtau).max())
    assert np.abs(eng.occ - r0["occ"]).max() < 1e-10 and np.abs(eng.densitymatrix - r0["dm"]).max() < 1e-10
    eng.dest_ctqmc()
