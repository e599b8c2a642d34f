"""CPU logic tests of the chain engine: the kernel source (w2dynamics_b200/csrc/engine_core.hpp) compiled with
W2B_HOST_EMU (one lane per chain, g++) is replayed move by move against the oracle.  This checks the *logic* the
CUDA build shares; the CUDA build itself is checked by the `gpu` tests through the product C ABI.  The host-emulation
library is test infrastructure only: the product module loads the CUDA library and nothing else."""
import numpy as np
import pytest

import _hostemu
from oracle.oracle import OracleChain
from w2dynamics_b200 import problems

from _parity import assert_replay                      # integer state bit-exact; det / trace ratios 1e-10 (tests/_parity.py)


def make(p, nchains=4, nbins=2, seed=43890):
    eng = _hostemu.make()
    eng.set_option("nchains", nchains)
    eng.set_option("nbins", nbins)
    eng.init_rnd(seed)
    eng.init_paras(p.pstring())
    eng.init_solver(p.u_matrix, p.ftau, p.muimp, None, 1)
    eng.set_taudiffmax(p.qmc["TaudiffMax"])
    eng.init_counters()
    return eng


@pytest.mark.parametrize("cfg,nsteps,chain", [("C1", 20000, 0), ("C2", 20000, 3), ("C5", 6000, 1), ("C3", 20000, 2), ("C4", 15000, 1)])
def test_replay_parity(built, cfg, nsteps, chain):
    p = problems.make_problem(cfg, nftau=2001)
    eng = make(p)
    log = eng.debug_steps(chain, nsteps)
    orc = OracleChain(p, rng="philox", seed=43890, chain=chain)
    orc.init_counters()
    ref = orc.steps(nsteps, log=True)
    assert_replay(log, ref)
    assert set(np.unique(log[:, 0])) == {1.0, 2.0, 3.0, 4.0}             # all four move types were exercised
    # chain state after the replay: trace, determinant, sign, inverse matrices
    s, o = eng.debug_chain_scalars(chain), orc.scalars()
    assert s["error"] == 0 and orc.error == 0
    assert abs(s["trace_log"] - o["trace_log"]) < 1e-8 and abs(s["det_log"] - o["det_log"]) < 1e-8
    assert (s["trace_sign"], s["det_sign"], s["cfgsign"]) == (o["trace_sign"], o["det_sign"], o["cfgsign"])
    for f in range(2 * p.nbands):
        a, b = eng.debug_minv(chain, f // 2 + 1, f % 2 + 1), orc.minv(f)
        assert a.shape == b.shape
        if a.size:
            assert np.abs(a - b).max() < 1e-7 * max(1.0, np.abs(b).max())
    eng.dest_ctqmc()


def test_config_roundtrip_and_measurement(built):
    """get_mc_config -> oracle.set_config gives the same derived state; one measurement of the same
    configuration adds the same numbers to the accumulators (MeasGtauRem, histograms, sign, density matrix)."""
    p = problems.make_problem("C2", nftau=2001, Ntau=200, NLegMax=30)
    eng = make(p, nchains=1, nbins=1)
    eng.ctqmc_warmup(6000)
    n, osst, _ = eng.get_mc_config_scalars(0)
    taus, orbs, spins, cas, hyb, osst2, _ = eng.get_mc_config(n, 0)
    assert osst == osst2 and np.all(np.diff(taus) > 0) and np.all(hyb == -1)
    orc = OracleChain(p, seed=1, Ntau=200, NLegMax=30)
    orc.set_config(taus, orbs, spins, cas, osst)
    s, o = eng.debug_chain_scalars(0), orc.scalars()
    assert abs(s["trace_log"] - o["trace_log"]) < 1e-9 and abs(s["det_log"] - o["det_log"]) < 1e-8
    assert s["cfgsign"] == o["cfgsign"]
    # set_mc_config on the engine side reproduces its own state from scratch (LU inverse)
    before = [eng.debug_minv(0, f // 2 + 1, f % 2 + 1) for f in range(4)]
    eng.set_mc_config(n, taus, orbs, spins, cas, hyb, osst, 1, chain=0)
    s2 = eng.debug_chain_scalars(0)
    assert abs(s2["det_log"] - s["det_log"]) < 1e-8 and s2["cfgsign"] == s["cfgsign"] and s2["error"] == 0
    for f in range(4):
        a = eng.debug_minv(0, f // 2 + 1, f % 2 + 1)
        assert np.abs(a - before[f]).max() < 1e-8 * max(1.0, np.abs(a).max())
    # one measurement each
    eng.block(0, True)
    eng._ck(eng._load().w2b_finalize())
    orc.measure(0, reset=True)
    orc.measure_once()
    r = orc.results(want_rho2=True)
    assert np.abs(eng.gtau - r["gtau"]).max() < 1e-9 * max(1.0, np.abs(r["gtau"]).max())
    assert np.abs(eng.gleg - r["gleg"]).max() < 1e-9 * max(1.0, np.abs(r["gleg"]).max())
    assert np.array_equal(eng.histo, r["hist"])
    assert eng.mean_sign == r["sign"]
    assert np.abs(eng.occ - r["occ"]).max() < 1e-12
    assert np.abs(eng.rho1 - r["rho1"]).max() < 1e-12
    assert np.abs(eng.rho2 - r["rho2"]).max() < 1e-12                                       # Meas_rho2_from_Densitymatrix CTQMC.F90:2945-3017
    eng.dest_ctqmc()


def test_statistical_agreement_c1(built):
    """Engine population vs oracle chain on C1: G(tau), occupation, <k>, sign agree within Monte-Carlo error bars;
    particle-hole symmetry pins the exact values (occ = 0.5, G(tau) = G(beta - tau), sign = 1)."""
    p = problems.make_problem("C1", nftau=2001, Nmeas=300, Ntau=100, NLegMax=20)
    eng = make(p, nchains=32, nbins=8)
    eng.ctqmc_warmup(3000)
    eng.ctqmc_measure(1, 1)
    orc = OracleChain(p, seed=1234, Ntau=100, NLegMax=20)
    orc.warmup(20000)
    orc.measure(9600)
    r = orc.results()
    bins = eng.bins("gtau")
    assert bins.shape[0] == 8
    err = bins.std(0, ddof=1) / np.sqrt(bins.shape[0])
    g = eng.gtau
    # literal restatement of cfgsign = -get_Sign (CTQMC.F90:3574; TTrace default -1, Trace.F90:129): a sign-problem-free
    # run reports mean_sign = -1 and every estimator divides it out (CTQMC.F90:8697, 8730)
    assert eng.mean_sign == -1.0 and r["sign"] == -1.0
    assert abs(eng.occ[0, 0, 0, 0] - 0.5) < 0.02 and abs(eng.occ[0, 1, 0, 1] - 0.5) < 0.02
    z = (g - r["gtau"]) / np.sqrt(err ** 2 + err.mean() ** 2 + 1e-12)
    assert np.abs(z).mean() < 2.0 and np.abs(z).max() < 6.0
    assert np.abs(g[0, 0] - g[0, 0, ::-1]).mean() < 4 * err.mean() + 1e-3
    k_eng = (eng.histo * np.arange(eng.histo.shape[2])[None, None, :]).sum(2)
    k_orc = (r["hist"] * np.arange(r["hist"].shape[2])[None, None, :]).sum(2)
    assert np.abs(k_eng - k_orc).max() < 0.15
    eng.dest_ctqmc()


def test_parameter_string_semantics(built):
    """Parameters.F90 quirks: prefix match on the first hit, commas -> blanks, scientific-notation integers;
    unsupported sectors fail loudly instead of silently running something else."""
    p = problems.make_problem("C1", nftau=101)
    eng = _hostemu.make()
    eng.init_rnd(1)
    with pytest.raises(RuntimeError):
        eng.init_paras(p.pstring(segment=1))
    with pytest.raises(RuntimeError):
        eng.init_paras(p.pstring(offdiag=1))
    eng.init_paras(p.pstring(Nmeas="1e2", NCorr=10))
    eng.set_option("nchains", 2)
    eng.init_solver(p.u_matrix, p.ftau, p.muimp, None, 1)
    with pytest.raises(RuntimeError):
        eng.ctqmc_measure(2, 1)                     # worm sectors are outside the path
    eng.ctqmc_measure(1, 1)
    assert eng.scalar("cnt_steps") == 2 * 100 * 10
    eng.dest_ctqmc()
    eng.dest_paras()


@pytest.mark.parametrize("cfg", ["C3", "C4"])
def test_blocked_trace_measurement(built, cfg):
    """Blocked local trace (d > 1): cached ket/bra evaluation == from-scratch oracle trace for many configurations, and the
    beta/2 density matrix, per-state histograms, G(tau), Legendre accumulators equal the oracle's on identical configurations."""
    p = problems.make_problem(cfg, nftau=2001, Ntau=100, NLegMax=20)
    eng = make(p, nchains=1, nbins=1)
    orc = OracleChain(p, seed=1, Ntau=100, NLegMax=20)
    orc.measure(0, reset=True)
    for _ in range(25):
        eng.ctqmc_warmup(400)
        n, osst, _ = eng.get_mc_config_scalars(0)
        taus, orbs, spins, cas, hyb, _, _ = eng.get_mc_config(n, 0)
        orc.set_config(taus, orbs, spins, cas, osst)
        s, o = eng.debug_chain_scalars(0), orc.scalars()
        assert abs(s["trace_log"] - o["trace_log"]) < 1e-9 and s["trace_sign"] == o["trace_sign"] and s["cfgsign"] == o["cfgsign"]
        eng.block(0, True)
        orc.measure_once()
    eng._ck(eng._load().w2b_finalize())
    r = orc.results(want_rho2=True)
    assert np.abs(r["rho2"]).max() > 0.1
    for a, b in ((eng.gtau, r["gtau"]), (eng.gleg, r["gleg"]), (eng.occ, r["occ"]), (eng.rho1, r["rho1"]), (eng.rho2, r["rho2"]),
                 (eng.densitymatrix, r["densitymatrix"]), (eng.outerstatehisto, r["hist-state"]),
                 (eng.tracecontribstates, r["contrib-state"]), (eng.tracecontribsuperstates, r["contrib-sst"])):
        assert np.abs(a - b).max() < 1e-9 * max(1e-3, np.abs(b).max())
    assert np.array_equal(eng.histo, r["hist"]) and eng.mean_sign == r["sign"]
    assert abs(np.trace(eng.densitymatrix) - 1.0) < 1e-9
    eng.dest_ctqmc()


@pytest.mark.parametrize("cfg,nsteps", [("C2", 4000), ("C4", 3000)])
def test_chain_independent_of_population(built, cfg, nsteps):
    """Scheduler invariance: a chain's move sequence (integer state, ratios, RNG draws) does not depend on how many other
    chains share the device -- work queues, rounds, rare-move hand-over and ticket order all change with the population, the
    per-chain Markov chain must not.  (1x1 blocks: CTA phase scheduler; blocked traces: propose / trace / decide rounds.)"""
    p = problems.make_problem(cfg, nftau=2001)
    logs = []
    for nchains in (2, 7):
        eng = make(p, nchains=nchains)
        logs.append(eng.debug_steps(1, nsteps))
        assert eng.scalar("error_flags") == 0
        eng.dest_ctqmc()
    assert np.array_equal(logs[0], logs[1])


@pytest.mark.parametrize("cfg,nsteps,chain", [("C2", 12000, 3), ("C5", 3000, 1)])
def test_round_scheduler_on_scalar_blocks(built, cfg, nsteps, chain):
    """The round kernels (propose / decide + rare stream) are the default for blocked traces only, but they are a complete
    scheduler: with 1x1 blocks the rare stream runs whole global moves / tau shifts.  Same replay parity against the oracle."""
    p = problems.make_problem(cfg, nftau=2001)
    eng = _hostemu.make()
    eng.set_option("nchains", 4); eng.set_option("nbins", 2); eng.set_option("scheduler", 2)
    eng.init_rnd(43890); eng.init_paras(p.pstring()); eng.init_solver(p.u_matrix, p.ftau, p.muimp, None, 1)
    eng.set_taudiffmax(p.qmc["TaudiffMax"]); eng.init_counters()
    log = eng.debug_steps(chain, nsteps)
    orc = OracleChain(p, rng="philox", seed=43890, chain=chain)
    orc.init_counters()
    assert_replay(log, orc.steps(nsteps, log=True))
    assert eng.scalar("rounds") > 0 and eng.scalar("error_flags") == 0
    eng.dest_ctqmc()


@pytest.mark.parametrize("cfg,nsteps,chain,pouter", [("C5", 5000, 1, 0.0), ("C2", 8000, 3, 0.0), ("C1", 6000, 0, 0.2)])
def test_autonomous_scheduler(built, cfg, nsteps, chain, pouter):
    """Scheduler 3 (large orders with 1x1 blocks: one warp takes a chain through all its moves, no CTA phases; engine_core.hpp
    `chain_sweep`): the same per-chain move sequence, so the same replay parity against the oracle, outer pair moves included."""
    p = problems.make_problem(cfg, nftau=2001, PercentageOuterMove=pouter)
    eng = _hostemu.make()
    eng.set_option("nchains", 4); eng.set_option("nbins", 2); eng.set_option("scheduler", 3)
    eng.init_rnd(43890); eng.init_paras(p.pstring()); eng.init_solver(p.u_matrix, p.ftau, p.muimp, None, 1)
    eng.set_taudiffmax(p.qmc["TaudiffMax"]); eng.init_counters()
    log = eng.debug_steps(chain, nsteps)
    orc = OracleChain(p, rng="philox", seed=43890, chain=chain)
    orc.init_counters()
    assert_replay(log, orc.steps(nsteps, log=True))
    assert eng.scalar("error_flags") == 0
    eng.dest_ctqmc()


def test_capacity_overflow_is_loud(built):
    """A chain that outgrows its operator capacity raises a flag (never silent truncation); the warm-up that hit the limit fails
    right away (not after the whole run) and ctqmc_measure refuses to hand out results -- the counterpart of the reference's hard
    `stop`s."""
    p = problems.make_problem("C2", nftau=2001, Nmeas=5)
    eng = _hostemu.make()
    eng.set_option("nchains", 2); eng.set_option("nbins", 1); eng.set_option("ops_cap", 32)
    eng.init_rnd(43890); eng.init_paras(p.pstring()); eng.init_solver(p.u_matrix, p.ftau, p.muimp, None, 1)
    eng.set_taudiffmax(p.qmc["TaudiffMax"]); eng.init_counters()
    with pytest.raises(RuntimeError, match="capacity exceeded during warm-up"):
        eng.ctqmc_warmup(20000)
    assert int(eng.scalar("error_flags")) & 1
    with pytest.raises(RuntimeError, match="capacity"):
        eng.ctqmc_measure(1, 1)
    eng.dest_ctqmc()


def _autocorracc(series, nlevels):
    """numpy restatement of autocorracc_add / autocorracc_blocks (Accumulators.F90:285-361) for series[n, comp]."""
    ncomp = series.shape[1]
    spill, mean, sq = (np.zeros((nlevels, ncomp)) for _ in range(3))
    for n0, val in enumerate(series):
        n, lvl, v = n0, 0, val
        while True:
            spill[lvl] += v
            incr1 = v - mean[lvl]; mean[lvl] += incr1 / (n + 1)
            sq[lvl] += incr1 * (v - mean[lvl])
            if not (n & 1) or lvl == nlevels - 1:
                break
            n //= 2
            v = spill[lvl] / 2; spill[lvl] = 0
            lvl += 1
    out, n = np.zeros((nlevels, ncomp)), series.shape[0]
    for lvl in range(nlevels):
        with np.errstate(divide="ignore", invalid="ignore"):
            out[lvl] = sq[lvl] / (n - 1) if n > 1 else sq[lvl] / 0.0
        n //= 2
    return out, mean[0]


def test_autocorrelation_accumulator(built):
    """AccumGtau = 2: logarithmic-binning variance of G(tau) per level.  The engine keeps one series per
    chain-group bin (group mean of the per-measurement contributions) and reports the single-chain equivalent
    (group variance x chains per bin, averaged over bins); checked against a numpy restatement of the reference accumulator
    fed with the very same per-measurement accumulator increments."""
    import ctypes as C
    nmeas, nbins, nchains, ntau = 21, 2, 6, 40

    def engine(accum):
        p = problems.make_problem("C1", nftau=2001, Nmeas=nmeas, Ntau=ntau, NLegMax=10, AccumGtau=accum)
        eng = make(p, nchains=nchains, nbins=nbins)
        eng.ctqmc_warmup(1500)
        return eng, p
    eng, p = engine(2)
    eng.ctqmc_measure(1, 1)
    blocks = eng.gtau_blocks
    nlev = blocks.shape[-1]
    assert nlev == 4 and blocks.shape[:3] == (1, 2, ntau)                 # ilog2(21) = 4 levels
    gtau_a = eng.gtau.copy()
    eng.dest_ctqmc()
    # the same Markov chains, one accumulation block at a time
    eng, p = engine(1)
    eng.init_counters()
    L = eng._load()
    n = C.c_size_t()
    eng._ck(L.w2b_read_accum(None, C.c_size_t(0), C.byref(n)))
    ncomp, series, prev = 2 * ntau, [], np.zeros((nbins, 2 * ntau))
    for _ in range(nmeas):
        eng.block(int(p.qmc.get("NCorr", 100)), True)
        buf = np.zeros(n.value)
        eng._ck(L.w2b_read_accum(buf.ctypes.data_as(C.POINTER(C.c_double)), C.c_size_t(n.value), C.byref(n)))
        cur = buf.reshape(nbins, -1)[:, :ncomp]                            # G(tau) accumulators lead every bin's slab: [flavour][tau]
        series.append((cur - prev) / (nchains / nbins)); prev = cur.copy()
    series = np.array(series)                                              # [n, bin, comp]
    ref = sum((nchains / nbins) * _autocorracc(series[:, b], nlev)[0] for b in range(nbins)) / nbins
    ref = ref.reshape(nlev, 2, ntau)                                       # [level][spin (one band)][tau]
    for lvl in range(nlev):
        assert np.allclose(blocks[0, :, :, lvl], ref[lvl], rtol=1e-12, atol=1e-300), lvl
    eng._ck(L.w2b_finalize())
    assert np.array_equal(eng.gtau, gtau_a)                                # the mean is unaffected by the accumulator choice
    eng.dest_ctqmc()


def test_operator_list_invariants_under_stress(built):
    """Counterpart of the reference's testsuite/ctqmc.tests/3-operator-list-stress-test.F90 (75 000 insert / remove / shift
    steps at beta / taudiff_max = 10, three orbitals; it walks the doubly linked operator list and asserts time ordering,
    per-flavour sub-list consistency and the pre-window counts).  The engine keeps packed time-sorted arrays instead of lists;
    the same invariants are checked on what it hands back: strictly increasing times inside (0, beta), as many creators as
    annihilators per flavour and both equal to the order of that flavour's inverse matrix, a superstate sequence that closes
    (the oracle rebuilds the configuration from scratch without error and finds the same trace, determinant and sign), and
    M * Delta = 1 for the from-scratch hybridisation matrix of the stored time order."""
    p = problems.make_problem("C3", nftau=2001)                      # beta 40, taudiff_max 4: 9 windows, window shifts every 16 attempts
    eng = make(p, nchains=2)
    orc = OracleChain(p, seed=1)
    seen_orders = set()
    for _ in range(30):
        eng.ctqmc_warmup(2500)
        for chain in range(2):
            n, osst, _ = eng.get_mc_config_scalars(chain)
            taus, orbs, spins, cas, hyb, osst2, _ = eng.get_mc_config(n, chain)
            assert osst == osst2 and n % 2 == 0
            assert np.all(np.diff(taus) > 0) and (n == 0 or (taus[0] > 0 and taus[-1] < p.beta))
            for b in range(p.nbands):
                for sp in range(2):
                    sel = (orbs == b + 1) & (spins == sp + 1)
                    kc, ka = int((cas[sel] == 1).sum()), int((cas[sel] == 2).sum())
                    m = eng.debug_minv(chain, b + 1, sp + 1)
                    assert kc == ka == m.shape[0]
                    seen_orders.add(kc)
            orc.set_config(taus, orbs, spins, cas, osst)
            s, o = eng.debug_chain_scalars(chain), orc.scalars()
            assert orc.error == 0 and s["error"] == 0
            assert abs(s["trace_log"] - o["trace_log"]) < 1e-8 and abs(s["det_log"] - o["det_log"]) < 1e-7
            assert (s["trace_sign"], s["det_sign"], s["cfgsign"]) == (o["trace_sign"], o["det_sign"], o["cfgsign"])
            for f in range(2 * p.nbands):
                a, c = eng.debug_minv(chain, f // 2 + 1, f % 2 + 1), orc.minv(f, scratch=True)
                if a.size:
                    assert np.abs(a - c).max() < 1e-6 * max(1.0, np.abs(c).max())
    assert len(seen_orders) > 5 and eng.scalar("cnt_acc_shift") > 0 and eng.scalar("cnt_acc_glob") >= 0
    eng.dest_ctqmc()
