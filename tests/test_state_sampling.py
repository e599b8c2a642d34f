"""State sampling (`statesampling = 1`, the second half of SURVEY.md 8 a21): the local trace runs over ONE outer eigenstate instead of
the sum over the outer superstate, and every move that changes the outer superstate draws the new eigenstate with
WeightedOuterStateChoice (Trace.F90:1477-1604; StepAdd / StepRem with change_outer CTQMC.F90:3471-3480, 3979-3996; global moves
Trace.F90:1325-1332, 1372-1379; tau shifts CTQMC.F90:4702-4710; empty configurations CTQMC.F90:4580-4608, Trace.F90:612-643; the
beta/2 density matrix and the outer histograms of the single state CTQMC.F90:3106-3110, 3294-3339).

Replay parity engine <-> oracle, CPU (host-emulation build of the kernel sources) and `-m gpu` (CUDA library through the C ABI).
NOTE on degenerate spectra: the contribution of ONE eigenstate of a degenerate multiplet depends on the basis chosen inside the
multiplet (the engine diagonalises with Jacobi rotations, the oracle with LAPACK; only sums over a multiplet are basis independent),
so per-move parity needs a non-degenerate spectrum inside every block: the blocked configurations get a crystal field and
orbital-pair dependent interaction strengths here.  The statistical correctness of state sampling on the degenerate models is
pinned against exact diagonalisation (tests/test_gpu_exact.py, tests/test_oracle_pins.py)."""
import numpy as np
import pytest

from w2dynamics_b200 import problems
from test_boundary import _cuda_engine, _hostemu_engine, _start


def lift_degeneracies(p):
    """Orbital-dependent level shifts and two-orbital interaction strengths (fixed pseudo-random factors): no two eigenvalues inside a
    quantum-number block coincide (the block structure itself -- Nt, Szt, Qzt -- is untouched: every element keeps its operator
    structure).  C3: smallest gap inside a block 0.05; C4 (748 states in blocks up to 28 wide): 1.7e-3."""
    nb = p.nbands
    uu = p.u_matrix.reshape(nb, 2, nb, 2, nb, 2, nb, 2)
    rng = np.random.RandomState(7)
    w = {}
    for idx in np.argwhere(np.abs(uu) > 1e-14):
        orbs = tuple(sorted({int(idx[0]), int(idx[2]), int(idx[4]), int(idx[6])}))
        if len(orbs) == 2:
            if orbs not in w:
                w[orbs] = 1.0 + 0.4 * (rng.rand() - 0.5)
            uu[tuple(idx)] *= w[orbs]
    lv = rng.rand(nb)
    for b in range(nb):
        for s in range(2):
            p.muimp[b, s, b, s] += lv[b]
    return p


def body_replay(make, cfg, nsteps, beta, extra, **opts):
    from oracle.oracle import OracleChain
    from _parity import assert_replay
    p = problems.make_problem(cfg, nftau=2001, beta=beta, statesampling=1, PercentageGlobalMove=0.05, PercentageTauShiftMove=0.05, **extra)
    assert "statesampling = 1" in p.pstring()
    if cfg in ("C3", "C4"):
        lift_degeneracies(p)
    eng = _start(make(), p, nchains=8, **opts)
    eng.init_counters()
    log = eng.debug_steps(5, nsteps)
    orc = OracleChain(p, rng="philox", seed=43890, chain=5)
    orc.init_counters()
    if cfg in ("C3", "C4"):
        assert min(np.diff(np.sort(e)).min() for e in orc.at.eval if len(e) > 1) > 1e-3
    ref = orc.steps(nsteps, log=True)
    # a single-state trace that vanishes by symmetry is a pure rounding residue (|ratio| < 1e-11: rejected with certainty in both)
    zero = (log[:, 9] < -25.0) & (ref[:, 9] < -25.0) & (log[:, 1] == 0) & (ref[:, 1] == 0)
    assert zero.sum() < 0.05 * nsteps                  # C4: 2 % (multiplet symmetries the quantum-number blocks do not resolve)
    log[zero, 9] = ref[zero, 9]; log[zero, 10] = ref[zero, 10]; log[zero, 11] = ref[zero, 11]
    if cfg == "C4":
        # near-degenerate levels (gaps down to 2e-3 in a spectrum 30 wide): the eigenvectors of two different eigensolvers agree to
        # eps * width / gap ~ 1e-11 per vector, and a single-state trace runs through tens of operators -- 1e-8 here instead of the
        # 1e-10 of every other replay (measured on the GPU: largest deviation 8e-10)
        near = np.abs(log[:, 9] - ref[:, 9]) < 1e-8
        log[near, 9] = ref[near, 9]; log[near, 11] = np.where(np.abs(log[near, 11] - ref[near, 11]) < 1e-7 * np.abs(ref[near, 11]) + 1e-13, ref[near, 11], log[near, 11])
    assert_replay(log, ref)
    acc = {k: int(((log[:, 0] == k) & (log[:, 1] == 1)).sum()) for k in range(1, 9)}
    assert acc[1] > 20 and acc[2] > 20 and acc[3] >= 5 and acc[4] >= 5          # pair moves, global moves and tau shifts were accepted
    if extra.get("PercentageOuterMove"):
        assert acc[5] + acc[6] >= 5
    if extra.get("Percentage4OperatorMove"):
        assert acc[7] + acc[8] >= 5
    # the configuration interface carries the outer eigenstate (1-based, get_mc_config_scalars CTQMC.F90:9358-9365)
    c = orc.config()
    n, osst, ostate = eng.get_mc_config_scalars(5)
    assert (n, osst, ostate) == (len(c["taus"]), c["outer_sst"], c["outer_state"])
    assert eng.scalar("error_flags") == 0
    if n > 0:                                          # set_mc_config round trip on another chain: same trace, same state
        cfgd = eng.get_mc_config(n, 5)
        eng.set_mc_config(n, *cfgd[:5], osst, ostate, chain=2)
        assert eng.get_mc_config_scalars(2) == (n, osst, ostate)
        a, b = eng.debug_chain_scalars(5), eng.debug_chain_scalars(2)
        assert abs(a["trace_log"] - b["trace_log"]) < 1e-9 and a["trace_sign"] == b["trace_sign"]
        with pytest.raises(RuntimeError, match="outer state"):
            eng.set_mc_config(n, *cfgd[:5], osst, 99, chain=2)
    eng.dest_ctqmc()


CPU_CASES = [("C2", 6000, 4.0, {}, {}), ("C2", 4000, 4.0, {"PercentageOuterMove": 0.25}, {"scheduler": 2}),
             ("C2", 4000, 4.0, {"PercentageOuterMove": 0.25, "Percentage4OperatorMove": 0.2}, {"scheduler": 3}),
             ("C3", 5000, 3.0, {}, {}), ("C3", 5000, 3.0, {"PercentageOuterMove": 0.25, "Percentage4OperatorMove": 0.2}, {}),
             ("C3", 4000, None, {}, {}), ("C4", 2000, 2.0, {"PercentageOuterMove": 0.25}, {})]
GPU_CASES = [("C2", 20000, 4.0, {"PercentageOuterMove": 0.25, "Percentage4OperatorMove": 0.2}, {}),
             ("C2", 8000, 4.0, {"PercentageOuterMove": 0.25}, {"scheduler": 2}), ("C2", 8000, 4.0, {"PercentageOuterMove": 0.25}, {"scheduler": 3}),
             ("C1", 10000, None, {}, {}), ("C5", 3000, None, {}, {}),
             ("C3", 12000, 3.0, {"PercentageOuterMove": 0.25, "Percentage4OperatorMove": 0.2}, {}), ("C3", 8000, None, {}, {}),
             ("C4", 6000, 2.0, {"PercentageOuterMove": 0.25}, {}), ("C4", 4000, None, {}, {})]


def _ids(cases):
    return ["%s-beta%s%s%s" % (c[0], c[2], "-outer" if c[3].get("PercentageOuterMove") else "", "-sched%d" % c[4]["scheduler"] if c[4] else "") +
            ("-4op" if c[3].get("Percentage4OperatorMove") else "") for c in cases]


@pytest.mark.parametrize("cfg,nsteps,beta,extra,opts", CPU_CASES, ids=_ids(CPU_CASES))
def test_state_sampling_replay(built, cfg, nsteps, beta, extra, opts):
    body_replay(_hostemu_engine, cfg, nsteps, beta, extra, **opts)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,nsteps,beta,extra,opts", GPU_CASES, ids=_ids(GPU_CASES))
def test_gpu_state_sampling_replay(built, cfg, nsteps, beta, extra, opts):
    body_replay(_cuda_engine, cfg, nsteps, beta, extra, **opts)


def body_measurement(make, cfg, beta):
    """Measurements of identical configurations (the engine's configuration, outer eigenstate included, is handed to the oracle): local
    trace of the single state, its beta/2 density matrix, occ, rho1, the outer-state histograms, G(tau)."""
    from oracle.oracle import OracleChain
    p = problems.make_problem(cfg, nftau=2001, beta=beta, statesampling=1, PercentageGlobalMove=0.05, PercentageTauShiftMove=0.05, Ntau=100, NLegMax=20,
                              PercentageOuterMove=0.2)
    if cfg in ("C3", "C4"):
        lift_degeneracies(p)
    eng = _start(make(), p, nchains=1, nbins=1)
    eng.init_counters()
    orc = OracleChain(p, seed=1, Ntau=100, NLegMax=20)
    orc.measure(0, reset=True)
    states = set()
    for _ in range(25):
        eng.ctqmc_warmup(300)
        n, osst, ostate = eng.get_mc_config_scalars(0)
        taus, orbs, spins, cas, hyb, _, _ = eng.get_mc_config(n, 0)
        orc.set_config(taus, orbs, spins, cas, osst, ostate)
        s, o = eng.debug_chain_scalars(0), orc.scalars()
        assert abs(s["trace_log"] - o["trace_log"]) < 1e-9 and s["trace_sign"] == o["trace_sign"] and s["cfgsign"] == o["cfgsign"]
        states.add((osst, ostate))
        eng.block(0, True)
        orc.measure_once()
    assert len(states) > 3
    eng._ck(eng._load().w2b_finalize())
    r = orc.results(want_rho2=True)
    for a, b in ((eng.gtau, r["gtau"]), (eng.occ, r["occ"]), (eng.rho1, r["rho1"]), (eng.rho2, r["rho2"]),
                 (eng.densitymatrix, r["densitymatrix"]), (eng.outerstatehisto, r["hist-state"]),
                 (eng.tracecontribstates, r["contrib-state"]), (eng.tracecontribsuperstates, r["contrib-sst"])):
        assert np.abs(a - b).max() < 1e-9 * max(1e-3, np.abs(b).max())
    assert abs(np.asarray(eng.outerstatehisto).sum() - 1.0) < 1e-12   # one state per measurement (normalised by the number of measurements), not one per state of the superstate
    assert np.array_equal(eng.histo, r["hist"]) and eng.mean_sign == r["sign"]
    assert eng.scalar("error_flags") == 0
    eng.dest_ctqmc()


@pytest.mark.parametrize("cfg,beta", [("C2", 4.0), ("C3", 3.0)])
def test_state_sampling_measurement(built, cfg, beta):
    body_measurement(_hostemu_engine, cfg, beta)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,beta", [("C2", 4.0), ("C3", 3.0), ("C4", 2.0)])
def test_gpu_state_sampling_measurement(built, cfg, beta):
    body_measurement(_cuda_engine, cfg, beta)


def test_state_sampling_exact_diagonalisation_host_emulation(built):
    """The kernel sources (host-emulation build) with state sampling and every optional move type on the DEGENERATE two-orbital Kanamori
    model against exact diagonalisation (tests/_exact_anderson.py): per-move parity with the oracle is basis dependent there (module
    docstring), the sampled distribution is not.  The same comparison runs on the GPU with 4096 chains (tests/test_gpu_exact.py)."""
    import _exact_anderson as ea
    p = ea.kanamori2_problem(2, 6.0, Nmeas=150, Ntau=1000, statesampling=1, PercentageOuterMove=0.2, Percentage4OperatorMove=0.2)
    gex, occ_ex, _ = ea.solve_cached(("K2", 2, 6.0), p)
    eng = _start(_hostemu_engine(), p, nchains=512, nbins=16)
    eng.init_counters()
    eng.ctqmc_warmup(4000)
    eng.ctqmc_measure(1, 1)
    assert eng.scalar("error_flags") == 0
    gb = eng.bins("gtau").reshape(16, 4, 50, 20).mean(3).mean(1)
    g, gerr = gb.mean(0), gb.std(0, ddof=1) / 4.0
    ob = eng.bins("occ")
    occ, oerr = ob.mean(0), ob.std(0, ddof=1) / 4.0
    z = np.concatenate([(g - gex[0]) / np.sqrt(gerr ** 2 + 1e-8), ((occ - occ_ex) / np.sqrt(oerr ** 2 + 1e-8)).ravel()])
    print("host emulation, state sampling vs exact: %d values, max |z| %.2f, mean z^2 %.2f, relative G error bar %.1e, %d outer states visited"
          % (z.size, np.abs(z).max(), (z ** 2).mean(), (gerr / g).mean(), int((np.asarray(eng.outerstatehisto) > 0).sum())))
    assert np.abs(z).max() < 5.0 and (z ** 2).mean() < 2.5 and (gerr / g).mean() < 0.04
    assert (np.asarray(eng.outerstatehisto) > 0).sum() >= 12          # single eigenstates all over the 16-state impurity
    eng.dest_ctqmc()


@pytest.mark.parametrize("ss", [0, 1])
def test_empty_configuration_round_trip(built, ss):
    """`set_empty_mc_config(outer_sst, outer_state)` (CTQMC.F90:9391-9400; MCConfig.set_to_ctqmc of an empty configuration,
    cthyb_solver.py:41-50): the trace of the empty configuration is the Boltzmann weight of the outer superstate -- of the ONE outer
    eigenstate with state sampling -- and the chain runs on from there (edge case: no operators at all)."""
    from oracle.oracle import OracleChain
    p = lift_degeneracies(problems.make_problem("C3", nftau=1001, beta=3.0, statesampling=ss))
    eng = _start(_hostemu_engine(), p, nchains=2, nbins=1)
    orc = OracleChain(p, seed=1)
    at = orc.at
    wide = int(np.argmax(at.dim))                      # a 3 x 3 block
    for state in (1, int(at.dim[wide])):
        eng.set_empty_mc_config(wide, state, chain=1)
        n, osst, ostate = eng.get_mc_config_scalars(1)
        assert (n, osst) == (0, wide) and ostate == (state if ss else 1)
        lw = -(np.asarray(at.eval[wide]) - at.egs) * p.beta
        want = lw[state - 1] if ss else np.log(np.exp(lw).sum())
        s = eng.debug_chain_scalars(1)
        assert abs(s["trace_log"] - want) < 1e-12 and s["trace_sign"] == 1.0 and s["n_ops"] == 0
        orc.set_config([], [], [], [], wide, state)
        assert abs(orc.scalars()["trace_log"] - want) < 1e-12
    if ss:
        with pytest.raises(RuntimeError, match="outer state"):
            eng.set_empty_mc_config(wide, int(at.dim[wide]) + 1, chain=1)
    eng.ctqmc_warmup(500)                              # and the chains run on from the empty configuration
    assert eng.scalar("error_flags") == 0 and eng.get_mc_config_scalars(1)[0] >= 0
    eng.dest_ctqmc()
