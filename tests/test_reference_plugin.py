"""Drop-in boundary test: the reference's OWN, unmodified solver plugin `w2dyn.dmft.cthyb_solver.CtHybSolver`
(imported from /root/reference) drives the engine through the `mctqmc` surface -- only the module behind
`from w2dyn.auxiliaries.CTQMC import mctqmc` (cthyb_solver.py:21) is swapped.  mpi4py is absent in this image, so a
single-rank stub communicator stands in for it.  Runs only where /root/reference exists (this container); uses the
host-emulation build of the kernel logic on CPU, the CUDA library when a GPU is visible."""
import os
import sys
import types

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "w2dyn")), reason="reference checkout not present")


class _Comm:
    def Get_rank(self): return 0
    def Get_size(self): return 1
    def Barrier(self): pass
    def bcast(self, x, root=0): return x
    def Allreduce(self, sendbuf, recvbuf, op=None): pass
    def allgather(self, x): return [x]


def _qmc_defaults():
    """[QMC] section of the reference's configspec with its defaults (configobj is absent here, so a tiny parser)."""
    import re
    out, on = {}, False
    for line in open(os.path.join(REF, "w2dyn", "auxiliaries", "configspec")):
        line = line.split("#")[0].strip()
        if line.startswith("["):
            on = line == "[QMC]"
            continue
        m = re.match(r"^(\w+)\s*=\s*(\w+)\((.*)\)$", line)
        if not (on and m):
            continue
        key, typ, args = m.groups()
        d = re.search(r"default\s*=\s*('[^']*'|\"[^\"]*\"|[^,)]*)", args)
        val = d.group(1).strip() if d else "None"
        if val == "None":
            out[key] = None
        elif typ == "integer":
            out[key] = int(val)
        elif typ == "float":
            out[key] = float(val)
        elif typ == "boolean":
            out[key] = val.strip("'\"") in ("True", "true", "1")
        else:
            out[key] = val.strip("'\"")
    return out


def _install_stubs(engine):
    mpi4py = types.ModuleType("mpi4py")
    MPI = types.ModuleType("mpi4py.MPI")
    MPI.IN_PLACE, MPI.SUM, MPI.COMM_WORLD = object(), object(), _Comm()
    mpi4py.MPI = MPI
    sys.modules.setdefault("mpi4py", mpi4py)
    sys.modules.setdefault("mpi4py.MPI", MPI)
    shim = types.ModuleType("w2dyn.auxiliaries.CTQMC")
    shim.mctqmc = engine
    sys.modules["w2dyn.auxiliaries.CTQMC"] = shim
    if REF not in sys.path:
        sys.path.insert(0, REF)


def _engine():
    import torch
    if torch.cuda.is_available():
        from w2dynamics_b200.CTQMC import MCTQMC
        eng = MCTQMC()
    else:
        import _hostemu
        eng = _hostemu.make()
    eng.set_option("nchains", 16)
    eng.set_option("nbins", 4)
    return eng


def _c1_inputs(taudiffmax):
    """C1 problem in the plugin's input conventions + the [QMC] section."""
    from w2dynamics_b200 import problems
    p = problems.make_problem("C1", nftau=501)
    niw = 200
    qmc = _qmc_defaults()                              # every [QMC] key of the reference, at its default
    assert qmc["NCorr"] == 100 and qmc["Nftau"] == 5000 and "real_ham_tolerance" in qmc
    qmc.update(Nwarmups=2000, Nmeas=150, NCorr=50, Ntau=200, NLegMax=40, NLegOrder=40, Niw=niw, Nftau=501,
               TaudiffMax=p.qmc["TaudiffMax"] if taudiffmax == "fixed" else -1.0,   # -1 (the default): ctqmc_calibrate path
               AccumGtau=2 if taudiffmax == "fixed" else 1)                           # 2: autocorrelation accumulator -> "gtau-blocks"
    if taudiffmax == "giw":             # Z-space Matsubara estimator: Giw(NBands, 2, -Niw:Niw-1) has to match the problem's niw frequencies
        qmc.update(TaudiffMax=p.qmc["TaudiffMax"], MeasGiw=1, Niw=niw // 2)
    return p, qmc, niw


@pytest.mark.parametrize("taudiffmax", ["fixed", "calibrate", "giw"])
def test_restated_sequence_equals_plugin(built, taudiffmax):
    """tests/_plugin_sequence.py (what the `-m gpu` boundary tests drive the CUDA library with, /root/reference being absent on the
    GPU box) issues exactly the calls of the real, unmodified CtHybSolver: same subroutines, same order, same argument values,
    same module-variable reads and writes."""
    import _plugin_sequence as ps
    p, qmc, niw = _c1_inputs(taudiffmax)
    nb, beta = 1, p.beta
    # --- the real plugin against a recording proxy
    rec = ps.Recorder(_engine())
    _install_stubs(rec)
    import w2dyn.dmft.cthyb_solver as cs
    import w2dyn.dmft.impurity as imp
    import w2dyn.dmft.interaction as inter
    ftau5 = np.zeros((501, nb, 2, nb, 2))
    muimp = np.zeros((nb, 2, nb, 2))
    for s in range(2):
        ftau5[:, 0, s, 0, s] = p.ftau[0, s, 0, s]
        muimp[0, s, 0, s] = p.mu
    screening = np.zeros((501, nb, 2, nb, 2))
    interaction = inter.Density(nb, 2.0, 0.0, 0.0)
    interaction.quantum_numbers = ("auto",) + interaction.auto_quantum_numbers
    problem = imp.ImpurityProblem(beta, np.zeros((niw, nb, 2, nb, 2), complex), np.zeros((niw, nb, 2, nb, 2), complex), screening, ftau5, muimp, interaction)
    config = {"QMC": qmc, "General": {"FTType": "none" if taudiffmax == "giw" else "legendre", "beta": beta}, "Atoms": {"1": {"Nd": nb}}}
    solver = cs.CtHybSolver(config, 43890, 0, 0, 0, interactive=False, mpi_comm=_Comm())
    solver.set_problem(problem)
    cfg_a = []
    res_a = solver.solve(0, cfg_a)
    # --- the restatement against a recording proxy around a fresh engine
    rec2 = ps.Recorder(_engine())
    seq = ps.PluginSequence(rec2, qmc, 43890, interactive=False)
    seq.set_problem("Density", beta, nb, " ".join(interaction.auto_quantum_numbers), interaction.u_matrix, ftau5, muimp, screening)
    cfg_b = []
    res_b = seq.solve(0, cfg_b)
    a = [e for e in rec.log if e[1] != "shape"]
    b = [e for e in rec2.log if e[1] != "shape"]
    for i, (x, y) in enumerate(zip(a, b)):
        assert x == y, "call %d differs:\n  plugin      %r\n  restatement %r" % (i, x, y)
    assert len(a) == len(b)
    # same engine, same seed, same calls -> same numbers
    assert np.array_equal(res_a.other["gtau"].mean(), res_b["gtau"]) and np.array_equal(res_a.other["occ"].mean(), res_b["occ"])
    assert np.array_equal(cfg_a[0].taus, cfg_b[0].taus)
    for k in ("w2dyn.auxiliaries.CTQMC", "w2dyn.dmft.cthyb_solver"):
        sys.modules.pop(k, None)


@pytest.mark.parametrize("taudiffmax", ["fixed", "calibrate", "giw"])
def test_unmodified_cthybsolver_runs_on_engine(built, taudiffmax):
    from w2dynamics_b200 import problems
    eng = _engine()
    _install_stubs(eng)
    import w2dyn.dmft.cthyb_solver as cs
    import w2dyn.dmft.impurity as imp
    import w2dyn.dmft.interaction as inter
    import w2dyn.auxiliaries.transform as tf

    p, qmc, niw = _c1_inputs(taudiffmax)
    nb, beta = 1, p.beta
    # "giw": MeasGiw = 1 with FTType = none -- the plugin takes G(iw) = -ctqmc.giw (cthyb_solver.py:813-814, 989-990), no transform
    config = {"QMC": qmc, "General": {"FTType": "none" if taudiffmax == "giw" else "legendre", "beta": beta}, "Atoms": {"1": {"Nd": nb}}}
    iw = tf.matfreq(beta, "fermi", niw)
    tau = np.linspace(0, beta, 501)
    fiw, ftau = tf.hybr_from_sites(p.eps_bath, p.v_bath[:, None], iw, tau, beta)       # (nb, nb, niw), (nb, nb, ntau)
    fiw5 = np.zeros((niw, nb, 2, nb, 2), complex)
    ftau5 = np.zeros((501, nb, 2, nb, 2))
    for s in range(2):
        fiw5[:, 0, s, 0, s] = fiw[0, 0]
        ftau5[:, 0, s, 0, s] = ftau[0, 0].real
    muimp = np.zeros((nb, 2, nb, 2))
    for s in range(2):
        muimp[0, s, 0, s] = p.mu                       # lattice convention; the plugin negates it (cthyb_solver.py:255)
    g0inv = np.zeros_like(fiw5)
    for s in range(2):
        g0inv[:, 0, s, 0, s] = 1j * iw + p.mu - fiw[0, 0]
    interaction = inter.Density(nb, 2.0, 0.0, 0.0)
    interaction.quantum_numbers = ("auto",) + interaction.auto_quantum_numbers        # as config.interaction_from_cfg does (config.py:272)
    problem = imp.ImpurityProblem(beta, g0inv, fiw5, np.zeros((2, nb, 2, nb, 2)), ftau5, muimp, interaction)

    solver = cs.CtHybSolver(config, 43890, 0, 0, 0, interactive=False, mpi_comm=_Comm())
    solver_tdm = []
    orig = eng.set_taudiffmax
    eng.set_taudiffmax = lambda t: (solver_tdm.append(float(t)), orig(t))[1]
    solver.set_problem(problem)
    mc_cfg = []
    result = solver.solve(0, mc_cfg)
    if taudiffmax == "calibrate":       # calibrate_taudiff_max (cthyb_solver.py:899-944) ran both phases and set a window
        assert 0.0 < solver_tdm[0] <= beta

    giw = result.giw.mean()
    assert giw.shape == (niw, nb, 2, nb, 2) and np.all(np.isfinite(giw))
    occ = result.other["occ"].mean()
    assert abs(occ[0, 0, 0, 0] - 0.5) < 0.06 and abs(occ[0, 1, 0, 1] - 0.5) < 0.06       # half filling
    assert float(result.other["sign"].mean()) == -1.0
    gtau = result.other["gtau"].mean()
    assert gtau.shape == (nb, 2, 200) and gtau.min() > 0
    # G(iw) from the Legendre coefficients: particle-hole symmetric problem -> purely imaginary, negative imaginary part
    g = giw[niw // 2:niw // 2 + 10, 0, 0, 0, 0]
    assert np.all(g.imag < 0) and np.abs(g.real).max() < 0.08
    if taudiffmax == "giw":             # the measured G(iw) is the Fourier integral of the measured G(tau) (200 bins: 3 % at the lowest frequencies)
        assert "giw-meas" in result.other and result.other["giw-meas"].mean().shape == (nb, 2, niw)
        tmid = (np.arange(200) + 0.5) * beta / 200
        ft = -(gtau[0, 0][None, :] * np.exp(1j * iw[niw // 2:niw // 2 + 5, None] * tmid[None, :])).sum(1) * beta / 200
        assert np.abs(ft - g[:5]).max() < 0.03 * np.abs(g[:5]).max()
    if taudiffmax == "fixed":           # cthyb_solver.py:1024-1026 reads ctqmc.gtau_blocks; ilog2(150) = 7 binning levels
        blocks = result.other["gtau-blocks"].mean()
        assert blocks.shape == (nb, 2, 200, 7) and np.all(blocks[..., 0] >= 0) and np.all(np.isfinite(blocks[..., :5]))
        import w2dyn.auxiliaries.postprocessing as pp
        tauint = pp.gtau_tauint(blocks)                  # the reference's own logarithmic-binning analysis accepts it
        assert tauint.shape == (nb, 2, 200)
    # ReuseMCConfig hook: the plugin fetched the configuration through get_mc_config
    assert len(mc_cfg) == 1 and mc_cfg[0].taus.shape == mc_cfg[0].orbs.shape
    for k in ("w2dyn.auxiliaries.CTQMC", "w2dyn.dmft.cthyb_solver"):
        sys.modules.pop(k, None)


def test_chain_group_bins_feed_distributedsample(built):
    """the reference's OWN `DistributedSample` (statistics.py:36-63, 99-132, 180-209) accepts n samples per
    rank.  The engine hands back `nbins` chain-group means (`mctqmc.bins(name)`), so `mean()` reproduces the population result and
    `stderr()` is the standard error over chain groups -- what the 3-line change at cthyb_solver.py:366,389-390 documented in
    INTEGRATION.md passes on.  Checked here with the unmodified class on gtau, gleg, occ and sign."""
    from w2dynamics_b200 import problems
    eng = _engine()
    eng.set_option("nchains", 32)
    eng.set_option("nbins", 8)
    _install_stubs(eng)
    sys.modules.pop("w2dyn.auxiliaries.statistics", None)
    import w2dyn.auxiliaries.statistics as st
    p = problems.make_problem("C1", nftau=501, Nmeas=40, Nwarmups=2000)
    eng.init_rnd(43890)
    eng.init_paras(p.pstring())
    eng.init_solver(p.u_matrix, p.ftau, p.muimp, None, 1)
    eng.set_taudiffmax(p.qmc["TaudiffMax"])
    eng.ctqmc_warmup(2000)
    eng.init_counters()
    eng.ctqmc_measure(1, 1)
    comm = _Comm()
    for name in ("gtau", "gleg", "occ"):
        pop = np.array(getattr(eng, name))
        bins = eng.bins(name)
        assert bins.shape == (8,) + pop.shape
        ds = st.DistributedSample(bins, comm)
        assert ds.ntotal == 8
        # every bin holds the same number of chains and the sign is -1 in every bin (no sign problem): mean of bins = population
        assert np.allclose(ds.mean(), pop, rtol=1e-12, atol=1e-14)
        sem = ds.stderr()
        ref = bins.std(axis=0, ddof=1) / np.sqrt(8.0)
        assert np.allclose(sem, ref, rtol=1e-12, atol=1e-300)
        if name == "gtau":
            assert np.all(sem > 0) and np.all(sem < np.abs(pop).max())           # genuine scatter between chain groups
    sign = st.DistributedSample(eng.bins("sign"), comm)
    assert float(sign.mean()) == float(eng.mean_sign) == -1.0 and float(sign.stderr()) == 0.0
    # what the unmodified plugin does instead (one sample per rank): SEM over ranks only -- a single rank has none
    one = st.DistributedSample([np.array(eng.gtau)], comm)
    assert one.ntotal == 1
    eng.dest_ctqmc()
    for k in ("w2dyn.auxiliaries.CTQMC", "w2dyn.auxiliaries.statistics"):
        sys.modules.pop(k, None)
