"""Z-space Matsubara estimator (`MeasGiw = 1`, the estimator half of SURVEY.md 8 f4): Giw(b, s, n) = zScaling / beta * sum over
measurements and (annihilator, creator) pairs of -cfgsign M(c, a) exp(i w_n (tau_a - tau_c)), n = -Niw .. Niw - 1 (CTQMC.F90:569-575,
2400-2414, measure_giw_nfft 1015-1056, normalisation 8815-8817; read by the plugin as `ctqmc.giw`, cthyb_solver.py:990, 814).  The
reference evaluates the sums with NFFT; engine and oracle sum them directly (the engine with a phase recurrence per pair).

* measurement parity on identical configurations, engine <-> oracle (CPU host emulation and `-m gpu`);
* consistency with the G(tau) estimator of the same run: giw is the Fourier integral of the binned G(tau) up to the bin width;
* `-m gpu`: against the EXACT G(i w) of an interacting impurity + bath model (dense diagonalisation, no oracle involved)."""
import numpy as np
import pytest

from w2dynamics_b200 import problems
from test_boundary import _cuda_engine, _hostemu_engine, _start


def body_parity(make, cfg, beta, niw):
    from oracle.oracle import OracleChain
    p = problems.make_problem(cfg, nftau=2001, beta=beta, MeasGiw=1, Niw=niw, Ntau=400, NLegMax=20)
    assert "MeasGiw = 1" in p.pstring() and ("Niw = %d" % niw) in p.pstring()
    eng = _start(make(), p, nchains=3, nbins=1)
    eng.init_counters()
    orc = OracleChain(p, seed=1, Ntau=400, NLegMax=20)
    orc.measure(0, reset=True)
    for it in range(12):
        eng.ctqmc_warmup(500)
        for chain in range(3):
            n, osst, ostate = eng.get_mc_config_scalars(chain)
            taus, orbs, spins, cas, hyb, _, _ = eng.get_mc_config(n, chain)
            orc.set_config(taus, orbs, spins, cas, osst)
            orc.measure_once()
        eng.block(0, True)
    eng._ck(eng._load().w2b_finalize())
    r = orc.results()
    g, go = np.asarray(eng.giw), np.asarray(r["giw"])
    assert g.shape == (p.nbands, 2, 2 * niw) and g.dtype == complex and go.shape == g.shape
    assert np.abs(go).max() > 0.05
    assert np.abs(g - go).max() < 1e-9 * np.abs(go).max()
    assert np.abs(np.asarray(eng.gtau) - r["gtau"]).max() < 1e-9 * np.abs(r["gtau"]).max()
    # G(-i w) = conj G(i w) for a real G(tau): index m <-> 2 Niw - 1 - m
    assert np.abs(g - np.conj(g[:, :, ::-1])).max() < 1e-12 * np.abs(g).max()
    # Fourier integral of the binned G(tau) of the same run (every pair sits in its bin: phase error <= w_n * bin width / 2)
    ntau = 400
    tmid = (np.arange(ntau) + 0.5) * p.beta / ntau
    wn = (2.0 * np.arange(-niw, niw) + 1.0) * np.pi / p.beta
    ft = (np.asarray(eng.gtau)[:, :, None, :] * np.exp(1j * wn[None, None, :, None] * tmid[None, None, None, :])).sum(-1) * p.beta / ntau
    low = np.abs(wn) * p.beta / ntau < 0.05
    assert np.abs(ft - g)[:, :, low].max() < 0.05 * np.abs(g).max()
    b = eng.bins("giw")
    assert b.shape == (1,) + g.shape and np.abs(b[0] - g).max() < 1e-12 * np.abs(g).max()
    assert eng.scalar("error_flags") == 0
    eng.dest_ctqmc()


@pytest.mark.parametrize("cfg,beta,niw", [("C1", None, 40), ("C2", 4.0, 37), ("C3", 3.0, 20)])
def test_giw_measurement_parity(built, cfg, beta, niw):
    body_parity(_hostemu_engine, cfg, beta, niw)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,beta,niw", [("C1", None, 300), ("C2", None, 100), ("C4", None, 50), ("C5", None, 20)])
def test_gpu_giw_measurement_parity(built, cfg, beta, niw):
    body_parity(_cuda_engine, cfg, beta, niw)


@pytest.mark.gpu
def test_gpu_giw_against_exact_diagonalisation(built):
    """Two-orbital full Kanamori model + two bath levels per flavour: the measured Giw of 4096 chains inside its chain-group error bars
    of the exact Matsubara Green's function (Lehmann representation of the dense diagonalisation, tests/_exact_anderson.py)."""
    import _exact_anderson as ea
    niw = 30
    p = ea.kanamori2_problem(2, 6.0, Nmeas=100, Ntau=1000, MeasGiw=1, Niw=niw)
    gex = ea.solve(p, niw=niw)[3][0]
    eng = _start(_cuda_engine(), p, nchains=4096, nbins=32)
    eng.init_counters()
    eng.ctqmc_warmup(6000)
    eng.ctqmc_measure(1, 1)
    assert eng.scalar("error_flags") == 0
    b = eng.bins("giw").mean(axis=(1, 2))                            # [bin, frequency], flavour average (all flavours are equivalent)
    g, gerr = b.mean(0), np.sqrt(b.real.var(0, ddof=1) + b.imag.var(0, ddof=1)) / np.sqrt(32.0)
    z = np.abs(g - gex) / np.sqrt(gerr ** 2 + 1e-10)
    print("Giw vs exact diagonalisation: %d frequencies, max |z| %.2f, mean z^2 %.2f, relative error bar %.1e" % (z.size, z.max(), (z ** 2).mean(), (gerr / np.abs(g)).mean()))
    assert z.max() < 5.0 and (z ** 2).mean() < 3.0 and (gerr / np.abs(g)).mean() < 0.02
    assert np.abs(np.asarray(eng.giw).mean(axis=(0, 1)) - g).max() < 1e-9
    eng.dest_ctqmc()


def test_oracle_giw_against_exact_diagonalisation(built):
    """The oracle's Matsubara estimator against the exact G(i w) of the two-orbital Kanamori model (16 chains): pins the sign, the
    frequency convention and the normalisation to something that is not a restatement."""
    import threading
    import _exact_anderson as ea
    from oracle.oracle import OracleChain
    niw = 20
    p = ea.kanamori2_problem(2, 6.0, NCorr=20, MeasGiw=1, Niw=niw)
    gex = ea.solve(p, niw=niw)[3][0]
    chains = [OracleChain(p, rng="philox", seed=4242, chain=7000 + i) for i in range(16)]

    def work(t):
        for c in chains[t::8]:
            c.warmup(4000)
            c.measure(6000)
    ths = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    gs = np.array([c.results()["giw"].mean(axis=(0, 1)) for c in chains])
    g, err = gs.mean(0), np.sqrt(gs.real.var(0, ddof=1) + gs.imag.var(0, ddof=1)) / 4.0
    z = np.abs(g - gex) / err
    print("oracle Giw vs exact: max |z| %.2f, mean z^2 %.2f, relative error bar %.1e" % (z.max(), (z ** 2).mean(), (err / np.abs(g)).mean()))
    assert z.max() < 5.0 and (z ** 2).mean() < 3.0 and (err / np.abs(g)).mean() < 0.06
