"""Boundary behaviour of the mctqmc surface beyond the happy path, on the host-emulation build (CPU) and -- same bodies, `-m gpu` --
on the CUDA library through the C ABI:
  * successive init_solver calls after ONE init_rnd draw fresh random numbers (the reference seeds once in CtHybSolver.__init__ and
    lets the generator run on, cthyb_solver.py:113), a new init_rnd reproduces the first solve;
  * `aborted` (register_trap CTQMC.F90:8076-8078, polls 8214 / 8430, truncation 8648-8658): 2 = signal before the first
    measurement, 1 = some measurements done, results normalised over those;
  * reuse_population: dest_ctqmc parks the population, the next init_solver revives it and rebuilds every chain from its operator
    list with the NEW hybridisation function (ReuseMCConfig for all chains, cthyb_solver.py:328-345 / Trace.F90:6872-7000);
  * index / range validation of the configuration entry points."""
import threading
import time

import numpy as np
import pytest

from w2dynamics_b200 import problems


def _hostemu_engine():
    import _hostemu
    return _hostemu.make()


def _cuda_engine():
    from w2dynamics_b200.CTQMC import MCTQMC
    e = MCTQMC()
    e.set_option("device", 0)
    return e


def _start(eng, p, nchains=8, nbins=4, seed=43890, init_rnd=True, **opts):
    eng.set_option("nchains", nchains)
    eng.set_option("nbins", nbins)
    for k, v in opts.items():
        eng.set_option(k, v)
    if init_rnd:
        eng.init_rnd(seed)
    eng.init_paras(p.pstring())
    eng.init_solver(p.u_matrix, p.ftau, p.muimp, None, 1)
    eng.set_taudiffmax(p.qmc["TaudiffMax"])
    return eng


def body_fresh_streams(make, bit_exact=True):
    p = problems.make_problem("C1", nftau=501, Nmeas=20, Ntau=50, NLegMax=10)
    eng = make()
    res = []
    for i, rnd in enumerate([True, False, False, True]):
        _start(eng, p, init_rnd=rnd)
        eng.ctqmc_warmup(300)
        eng.ctqmc_measure(1, 1)
        res.append(eng.gtau.copy())
        eng.dest_ctqmc()
    assert not np.array_equal(res[0], res[1]) and not np.array_equal(res[1], res[2]) and not np.array_equal(res[0], res[2])
    # init_rnd restarts the streams: replay tests stay reproducible.  The chains walk bit-identical paths; on the GPU the G(tau) bins are
    # summed with floating-point atomics from several warps, whose order is not fixed (seen once in 20 GPU runs: last-bit differences)
    if bit_exact:
        assert np.array_equal(res[0], res[3])
    else:
        assert np.abs(res[0] - res[3]).max() <= 1e-13 * np.abs(res[0]).max() and np.abs(res[0] - res[1]).max() > 1e-3 * np.abs(res[0]).max()


def body_aborted(make):
    p = problems.make_problem("C1", nftau=501, Nmeas=100000, Ntau=50, NLegMax=10)
    eng = _start(make(), p)
    L = eng._load()
    L.w2b_trap_notify.restype = None
    eng.ctqmc_warmup(200)
    assert int(eng.aborted) == 0
    L.w2b_trap_notify(10)                                       # SIGUSR1 before the first measurement
    eng.ctqmc_measure(1, 1)
    assert int(eng.aborted) == 2                                # cthyb_solver.py:347: the plugin raises KeyboardInterrupt
    eng.dest_ctqmc()
    # signal while sampling: the run stops between two measurements, the measurements taken so far are normalised
    eng = _start(make(), p)
    L = eng._load()
    eng.ctqmc_warmup(200)
    t = threading.Timer(0.5, lambda: L.w2b_trap_notify(15))
    t.start()
    t0 = time.time()
    eng.ctqmc_measure(1, 1)                                     # Nmeas = 100000 would take far longer than the timer
    t.join()
    assert int(eng.aborted) == 1 and 0 < eng.scalar("nmeas_done") < 100000 and time.time() - t0 < 120
    assert np.all(np.isfinite(eng.gtau)) and abs(eng.occ[0, 0, 0, 0] - 0.5) < 0.2 and float(eng.mean_sign) == -1.0
    eng.dest_ctqmc()
    # a real signal takes the same route: the traps are armed between init_solver and dest_ctqmc
    import os, signal
    eng = _start(make(), p)
    os.kill(os.getpid(), signal.SIGUSR1)
    eng.ctqmc_measure(1, 1)
    assert int(eng.aborted) == 2
    eng.dest_ctqmc()
    eng = _start(make(), problems.make_problem("C1", nftau=501, Nmeas=3, Ntau=50, NLegMax=10))
    eng.ctqmc_measure(1, 1)
    assert int(eng.aborted) == 0                                # the flag does not outlive the solve
    eng.dest_ctqmc()


def body_reuse_population(make, cfg):
    p = problems.make_problem(cfg, nftau=501, Nmeas=5, Ntau=50, NLegMax=10)
    eng = _start(make(), p, reuse_population=1)
    eng.ctqmc_warmup(1500)
    eng.ctqmc_measure(1, 1)
    before = [eng.get_mc_config(eng.get_mc_config_scalars(c)[0], c) if eng.get_mc_config_scalars(c)[0] else None for c in range(8)]
    assert any(b is not None for b in before)
    eng.dest_ctqmc()
    # next "DMFT iteration": same local problem, new hybridisation function
    p2 = problems.make_problem(cfg, nftau=501, Nmeas=5, Ntau=50, NLegMax=10)
    p2.ftau = p.ftau * 1.07
    _start(eng, p2, init_rnd=False, reuse_population=1)
    for c in range(8):
        n, osst, _ = eng.get_mc_config_scalars(c)
        if before[c] is None:
            assert n == 0
            continue
        now = eng.get_mc_config(n, c)
        assert n == before[c][0].size and np.array_equal(now[0], before[c][0]) and np.array_equal(now[1], before[c][1])
        # revived state == a from-scratch set_mc_config with the new F(tau)
        s = eng.debug_chain_scalars(c)
        eng.set_mc_config(n, *now[:5], osst, 1, chain=c)
        s2 = eng.debug_chain_scalars(c)
        assert s["error"] == 0 and s2["error"] == 0 and s["cfgsign"] == s2["cfgsign"]
        assert abs(s["det_log"] - s2["det_log"]) < 1e-9 and abs(s["trace_log"] - s2["trace_log"]) < 1e-9
    eng.ctqmc_warmup(0)
    eng.ctqmc_measure(1, 1)
    assert np.all(np.isfinite(eng.gtau)) and eng.scalar("error_flags") == 0
    eng.dest_ctqmc()
    # a different local problem is never revived
    p3 = problems.make_problem(cfg, nftau=501, Nmeas=5, Ntau=50, NLegMax=10)
    p3.muimp = p.muimp * 1.01
    _start(eng, p3, init_rnd=False, reuse_population=1)
    assert all(eng.get_mc_config_scalars(c)[0] == 0 for c in range(8))
    eng.set_option("reuse_population", 0)
    eng.dest_ctqmc()


def body_validation(make):
    p = problems.make_problem("C2", nftau=501)
    eng = _start(make(), p)
    with pytest.raises(RuntimeError, match="chain index"):
        eng.get_mc_config_scalars(8)
    with pytest.raises(RuntimeError, match="chain index"):
        eng.debug_chain_scalars(-1)
    t, o, s, ca, h = np.array([0.5, 1.0]), np.array([1, 1]), np.array([1, 1]), np.array([1, 2]), np.array([-1, -1])
    with pytest.raises(RuntimeError, match="outer superstate"):
        eng.set_mc_config(2, t, o, s, ca, h, 99, 1)
    with pytest.raises(RuntimeError, match="orbital"):
        eng.set_mc_config(2, t, np.array([3, 1]), s, ca, h, 0, 1)
    with pytest.raises(RuntimeError, match="tau outside"):
        eng.set_mc_config(2, np.array([0.5, 50.0]), o, s, ca, h, 0, 1)
    with pytest.raises(RuntimeError, match="ascending"):
        eng.set_mc_config(2, t[::-1].copy(), o, s, ca, h, 0, 1)
    eng.dest_ctqmc()
    # what is outside the implemented path is refused by name, never approximated
    for key, val in (("offdiag", 1), ("segment", 1), ("flavourchange_moves", 1), ("Percentage4OperatorMove", 1.5), ("PercentageOuterMove", -0.1)):
        q = problems.make_problem("C2", nftau=501, **{key: val})
        with pytest.raises(RuntimeError, match=key):
            make().init_paras(q.pstring())
    # 4-operator moves through the result surface: AccAdd4 / TryAdd4 (CTQMC.F90:8680-8695) -- NaN without them, as in the reference
    q = problems.make_problem("C2", nftau=501, beta=4.0, Nmeas=20, Percentage4OperatorMove=0.3)
    eng = _start(make(), q)
    eng.ctqmc_warmup(500); eng.ctqmc_measure(1, 1)
    assert 0.0 < float(eng.accadd4) < 1.0 and 0.0 < float(eng.accrem4) < 1.0 and np.isnan(float(eng.accflavc))
    eng.dest_ctqmc()
    eng = _start(make(), problems.make_problem("C2", nftau=501, Nmeas=5))
    eng.ctqmc_warmup(200); eng.ctqmc_measure(1, 1)
    assert np.isnan(float(eng.accadd4)) and np.isnan(float(eng.accrem4))
    eng.dest_ctqmc()


def body_calibrate_list(make):
    """ctqmc_calibrate phase 1 (CTQMC.F90:8118-8163, 3583-3586): `accpairtau` lists the tau distance of each of the first
    phase1pairnum accepted pair moves -- actual distances, not histogram bin edges."""
    p = problems.make_problem("C1", nftau=501, Ntau=50)
    eng = _start(make(), p)
    eng.set_taudiffmax(-1.0)
    eng.init_counters()
    eng.ctqmc_calibrate(True, 1000, 100, 1000, 100000)
    lst = eng.accpairtau
    assert lst.shape == (1000,) and np.all(lst > 0.0) and np.all(lst < p.beta)
    frac = np.abs(lst / (p.beta / 50) - np.rint(lst / (p.beta / 50)))
    assert (frac > 1e-6).mean() > 0.99 and len(np.unique(lst)) > 500          # a removal of a pair listed at its insertion repeats the distance
    q = np.sort(lst)[995]
    assert 0.2 * p.beta < q <= p.beta                             # the caller's 99.5 % quantile (cthyb_solver.py:906-909)
    eng.dest_ctqmc()


def body_symmetry_moves(make, cfg, nsteps):
    """User symmetry moves (NSymMove / SymMoveNN, symm_to_pstring cthyb_solver.py:769-774; gen_SymUpdate Trace.F90:1170-1182, the
    20 % branch of StepGlob CTQMC.F90:4463-4465): replay parity against the oracle with two declared orbital symmetries."""
    from oracle.oracle import OracleChain
    from _parity import assert_replay
    p = problems.make_problem(cfg, nftau=2001, PercentageGlobalMove=0.05)
    nb = p.nbands
    swap01 = list(range(2 * nb)); swap01[0], swap01[1] = 1, 0; swap01[nb], swap01[nb + 1] = nb + 1, nb      # orbitals 0 <-> 1, both spins
    cyc = [(b + 1) % nb + nb * s for s in range(2) for b in range(nb)]                                          # cyclic orbital shift
    p.symmetry_moves = [swap01, cyc]
    assert "NSymMove = 2" in p.pstring() and "SymMove02" in p.pstring()
    eng = _start(make(), p, nchains=8)
    eng.init_counters()
    log = eng.debug_steps(3, nsteps)
    orc = OracleChain(p, rng="philox", seed=43890, chain=3)
    orc.init_counters()
    ref = orc.steps(nsteps, log=True)
    assert_replay(log, ref)
    sym = log[(log[:, 0] == 3) & (log[:, 4] == 2)]
    assert sym.shape[0] > 5 and (sym[:, 1] == 1).any()            # symmetry moves were proposed, some accepted
    assert eng.scalar("error_flags") == 0
    eng.dest_ctqmc()
    with pytest.raises(RuntimeError, match="permutation"):
        q = problems.make_problem(cfg, nftau=501)
        q.symmetry_moves = [[0] * (2 * nb)]
        _start(make(), q)


def body_outer_moves(make, cfg, nsteps, beta=None, **opts):
    """Outer pair moves (`PercentageOuterMove > 0`: StepAdd / StepRem with change_outer, CTQMC.F90:3422-3449,
    3934-3961, 8441-8457): the superstates between the two operators are kept, the rest of the sequence is re-walked over
    tau = beta and the outer superstate changes.  Replay parity against the oracle (move types 5 / 6 in the log).  At the named
    configurations' temperatures these moves survive the quantum-number walk but are practically never accepted (a different outer
    superstate over a long beta); the high-temperature variants (`beta`) exercise the acceptance paths."""
    from oracle.oracle import OracleChain
    from _parity import assert_replay
    p = problems.make_problem(cfg, nftau=2001, beta=beta, PercentageOuterMove=0.25)
    assert "PercentageOuterMove = 0.25" in p.pstring()
    eng = _start(make(), p, nchains=8, **opts)
    eng.init_counters()
    log = eng.debug_steps(5, nsteps)
    orc = OracleChain(p, rng="philox", seed=43890, chain=5)
    orc.init_counters()
    ref = orc.steps(nsteps, log=True)
    assert_replay(log, ref)
    for kind in (5, 6):
        mv = log[log[:, 0] == kind]
        assert mv.shape[0] > 0.08 * nsteps and (mv[:, 1] == 2).any()
        if beta is not None:
            assert (mv[:, 1] == 1).sum() >= 10 and (mv[:, 1] == 0).sum() >= 10
    idx = np.nonzero((log[:, 0] >= 5) & (log[:, 1] == 1) & (np.arange(log.shape[0]) > 0))[0]
    assert (log[idx, 13] != log[idx - 1, 13]).all()                # an accepted outer move always changes the outer superstate
    assert eng.scalar("error_flags") == 0
    eng.dest_ctqmc()


def body_four_operator_moves(make, cfg, nsteps, beta=None, **opts):
    """4-operator moves with diagonal hybridisation (`Percentage4OperatorMove > 0`: StepAdd4 / StepRem4, CTQMC.F90:3615-3889,
    4153-4430, scheduler branch 8441-8457; the move half of SURVEY.md 8 f3): two pairs inserted / removed at once, determinant from
    scratch, proposal ratio summed over the orders of the inverse move.  Replay parity against the oracle (move types 7 / 8)."""
    from oracle.oracle import OracleChain
    from _parity import assert_replay
    p = problems.make_problem(cfg, nftau=2001, beta=beta, Percentage4OperatorMove=0.3)
    assert "Percentage4OperatorMove = 0.3" in p.pstring()
    eng = _start(make(), p, nchains=8, **opts)
    eng.init_counters()
    log = eng.debug_steps(5, nsteps)
    orc = OracleChain(p, rng="philox", seed=43890, chain=5)
    orc.init_counters()
    ref = orc.steps(nsteps, log=True)
    assert_replay(log, ref)
    for kind in (7, 8):
        mv = log[log[:, 0] == kind]
        assert mv.shape[0] > 0.1 * nsteps and (mv[:, 1] == 2).any() and (mv[:, 1] == 0).any()
        if beta is not None:
            assert (mv[:, 1] == 1).sum() >= 10
    acc = log[(log[:, 0] >= 7) & (log[:, 1] == 1)]
    idx = np.nonzero((log[:, 0] >= 7) & (log[:, 1] == 1) & (np.arange(log.shape[0]) > 0))[0]
    assert (np.abs(log[idx, 12] - log[idx - 1, 12]) == 4).all()     # an accepted 4-operator move changes the number of operators by four
    assert eng.scalar("error_flags") == 0
    # accepted 4-operator moves update the stored inverses by the rank-2 (two-step bordering / Schur-complement) update; the from-scratch
    # re-inversion of the reference (CTQMC.F90:3856-3866) only serves as the fall-back for a vanishing intermediate pivot
    n4 = eng.scalar("cnt_acc_add4") + eng.scalar("cnt_acc_rem4")
    if opts.get("four_rank2", 1):
        assert eng.scalar("cnt_four_rank2") >= 0.9 * n4 and (n4 == 0 or eng.scalar("cnt_four_rank2") > 0)
    else:
        assert eng.scalar("cnt_four_rank2") == 0
    # state after thousands of mixed moves: inverse matrices against the oracle's (fast updates on the engine's side, fast updates +
    # from-scratch rebuilds on the oracle's)
    for f in range(2 * p.nbands):
        a, b = eng.debug_minv(5, f // 2 + 1, f % 2 + 1), orc.minv(f)
        assert a.shape == b.shape and (a.size == 0 or np.abs(a - b).max() < 1e-8 * max(1.0, np.abs(b).max()))
    eng.dest_ctqmc()


# ------------------------------------------------------------------ CPU (host-emulation build of the same sources)
@pytest.mark.parametrize("cfg,nsteps,beta,opts", [("C2", 5000, None, {}), ("C3", 4000, None, {}), ("C3", 4000, 3.0, {}), ("C4", 2000, 2.0, {}),
                                                  ("C2", 4000, 4.0, {"scheduler": 2}), ("C2", 4000, 4.0, {"scheduler": 3}), ("C3", 4000, 3.0, {"four_rank2": 0})])
def test_four_operator_moves(built, cfg, nsteps, beta, opts):
    body_four_operator_moves(_hostemu_engine, cfg, nsteps, beta, **opts)


@pytest.mark.parametrize("cfg,nsteps,beta,opts", [("C2", 6000, None, {}), ("C2", 6000, 4.0, {}), ("C3", 5000, 3.0, {}), ("C4", 2500, 2.0, {}),
                                                  ("C2", 4000, 4.0, {"scheduler": 2})])
def test_outer_moves(built, cfg, nsteps, beta, opts):
    body_outer_moves(_hostemu_engine, cfg, nsteps, beta, **opts)


@pytest.mark.parametrize("cfg,nsteps", [("C2", 12000), ("C3", 6000)])
def test_symmetry_moves(built, cfg, nsteps):
    body_symmetry_moves(_hostemu_engine, cfg, nsteps)


def test_calibrate_phase1_list(built):
    body_calibrate_list(_hostemu_engine)



def test_fresh_streams_per_solve(built):
    body_fresh_streams(_hostemu_engine)


def test_aborted_flag(built):
    body_aborted(_hostemu_engine)


@pytest.mark.parametrize("cfg", ["C2", "C3"])
def test_reuse_population(built, cfg):
    body_reuse_population(_hostemu_engine, cfg)


def test_argument_validation(built):
    body_validation(_hostemu_engine)


# ------------------------------------------------------------------ GPU (CUDA library through the C ABI)
@pytest.mark.gpu
def test_gpu_fresh_streams_per_solve(built):
    body_fresh_streams(_cuda_engine, bit_exact=False)


@pytest.mark.gpu
def test_gpu_aborted_flag(built):
    body_aborted(_cuda_engine)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["C2", "C4"])
def test_gpu_reuse_population(built, cfg):
    body_reuse_population(_cuda_engine, cfg)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,nsteps", [("C2", 20000), ("C4", 12000)])
def test_gpu_symmetry_moves(built, cfg, nsteps):
    body_symmetry_moves(_cuda_engine, cfg, nsteps)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,nsteps,beta,opts", [("C2", 20000, None, {}), ("C4", 8000, None, {}), ("C5", 5000, None, {}), ("C1", 20000, None, {}),
                                                  ("C2", 20000, 4.0, {}), ("C3", 12000, 3.0, {}), ("C4", 8000, 2.0, {}), ("C2", 8000, 4.0, {"scheduler": 2})])
def test_gpu_outer_moves(built, cfg, nsteps, beta, opts):
    body_outer_moves(_cuda_engine, cfg, nsteps, beta, **opts)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,nsteps,beta,opts", [("C2", 20000, None, {}), ("C3", 12000, None, {}), ("C4", 6000, None, {}), ("C5", 4000, None, {}),
                                                  ("C3", 12000, 3.0, {}), ("C4", 6000, 2.0, {}), ("C2", 8000, 4.0, {"scheduler": 2}), ("C2", 8000, 4.0, {"scheduler": 3})])
def test_gpu_four_operator_moves(built, cfg, nsteps, beta, opts):
    body_four_operator_moves(_cuda_engine, cfg, nsteps, beta, **opts)


@pytest.mark.gpu
def test_gpu_calibrate_phase1_list(built):
    body_calibrate_list(_cuda_engine)


@pytest.mark.gpu
def test_gpu_argument_validation(built):
    body_validation(_cuda_engine)
