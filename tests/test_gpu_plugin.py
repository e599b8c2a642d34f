"""Drop-in boundary on the GPU: the reference plugin's call sequence (tests/_plugin_sequence.py -- proved call-for-call identical
to the unmodified `CtHybSolver` by tests/test_reference_plugin.py::test_restated_sequence_equals_plugin, which needs /root/reference
and therefore runs in the build container only) drives the CUDA library through the C ABI:
fixed window + `gtau-blocks`, the default `TaudiffMax = -1` two-phase `ctqmc_calibrate` path, and `ReuseMCConfig` round trips."""
import numpy as np
import pytest

from w2dynamics_b200 import problems

import _plugin_sequence as ps

pytestmark = pytest.mark.gpu


def _run(taudiffmax, cfg="C1", nchains=64, reuse=False, **over):
    from w2dynamics_b200.CTQMC import MCTQMC
    p = problems.make_problem(cfg, nftau=501)
    nb, beta = p.nbands, p.beta
    eng = MCTQMC()
    eng.set_option("nchains", nchains)
    eng.set_option("nbins", 8)
    eng.set_option("device", 0)
    qmc = ps.qmc_defaults()
    qmc.update(p.qmc)
    qmc.update(Nwarmups=2000, Nmeas=150, NCorr=50, Ntau=200, NLegMax=40, NLegOrder=40, Niw=200, Nftau=501,
               TaudiffMax=p.qmc["TaudiffMax"] if taudiffmax == "fixed" else -1.0, AccumGtau=2 if taudiffmax == "fixed" else 1,
               ReuseMCConfig=1 if reuse else 0)
    qmc.update(over)
    tdm = []
    orig = eng.set_taudiffmax
    eng.set_taudiffmax = lambda t: (tdm.append(float(t)), orig(t))[1]
    seq = ps.PluginSequence(eng, qmc, 43890)
    ftau5 = np.ascontiguousarray(p.ftau.transpose(4, 0, 1, 2, 3))
    screening = np.zeros_like(ftau5)
    seq.set_problem(p.hamiltonian, beta, nb, p.quantum_numbers, p.u_matrix, ftau5, -p.muimp, screening)
    return seq, eng, p, tdm


@pytest.mark.parametrize("taudiffmax", ["fixed", "calibrate"])
def test_plugin_sequence_on_cuda_library(built, taudiffmax):
    seq, eng, p, tdm = _run(taudiffmax)
    mc_cfg = []
    r = seq.solve(0, mc_cfg)
    assert eng._libpath.endswith("libw2b_cthyb.so")
    if taudiffmax == "calibrate":                         # calibrate_taudiff_max ran both phases and set a window (cthyb_solver.py:899-944)
        assert len(tdm) == 2 and 0.0 < tdm[0] <= p.beta and 0.0 < tdm[1]
    else:
        assert tdm == [p.qmc["TaudiffMax"]]
    occ = r["occ"]
    assert abs(occ[0, 0, 0, 0] - 0.5) < 0.04 and abs(occ[0, 1, 0, 1] - 0.5) < 0.04
    assert float(r["sign"]) == -1.0 and not seq.abort
    assert r["gtau"].shape == (1, 2, 200) and r["gtau"].min() > 0 and r["gleg"].shape == (1, 2, 40)
    assert r["rho2"].shape == (1, 2) * 4 and r["rho1"].shape == (1, 2) * 2 and r["densitymatrix"].shape == (4, 4)
    assert abs(np.trace(r["densitymatrix"]) - 1.0) < 1e-9
    assert r["accept-pair-tau"].dtype == np.int64 and r["accept-pair-tau"].sum() > 0
    assert r["time-sampling"] > 0 and r["time-simulation"] >= r["time-sampling"] and r["time-warmup"] > 0
    if taudiffmax == "fixed":                             # cthyb_solver.py:1024-1026: ilog2(150) = 7 binning levels
        assert r["gtau-blocks"].shape == (1, 2, 200, 7) and np.all(r["gtau-blocks"][..., 0] >= 0)
    assert len(mc_cfg) == 1 and mc_cfg[0].taus.shape == mc_cfg[0].orbs.shape


def test_reuse_mc_config_roundtrip(built):
    """ReuseMCConfig (cthyb_solver.py:328-345): the configuration handed back by one solve seeds the next one; a second solve with a
    short warm-up starts from it (set_mc_config rebuilds M by LU, trace and signs from scratch on the device)."""
    seq, eng, p, _ = _run("fixed", cfg="C2", reuse=True, Nwarmups2Plus=0, Nmeas=20)
    store = []
    seq.solve(0, store)
    assert len(store) == 1 and store[0].taus.size > 0
    first = store[0]
    seq2, eng2, _, _ = _run("fixed", cfg="C2", reuse=True, Nwarmups2Plus=0, Nmeas=1, NCorr=1)
    seen = {}
    orig = eng2.set_mc_config
    eng2.set_mc_config = lambda n, *a, **k: (seen.setdefault("n", n), orig(n, *a, **k))[1]
    orig_w = eng2.ctqmc_warmup

    def warm(n):
        seen["warm"] = n
        s = eng2.debug_chain_scalars(0)
        seen["n_ops"], seen["err"] = s["n_ops"], s["error"]
        orig_w(n)
    eng2.ctqmc_warmup = warm
    store2 = [first]
    seq2.solve(1, store2)
    assert seen["n"] == first.taus.size and seen["warm"] == 0 and seen["n_ops"] == first.taus.size and seen["err"] == 0
    assert len(store2) == 1
