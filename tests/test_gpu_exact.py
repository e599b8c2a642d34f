"""The CUDA engine against EXACT results of interacting Anderson impurity models (no oracle involved): the impurity of a named
configuration coupled to one to three bath levels per flavour is diagonalised densely in the full Fock space of impurity + bath
(tests/_exact_anderson.py, numpy only); G(tau) in 50 tau bins (exact bin averages), the Legendre coefficients l < 12 (closed form) and
all <n_i n_j> must lie inside the engine's own
chain-group error bars.  This pins the whole hot path -- proposals, window, quantum-number walk, determinant fast updates, local
trace (1x1 blocks AND the blocked DMMA path: two-orbital Kanamori model with spin-flip / pair-hopping terms), signs, G(tau) and density-matrix
estimators, normalisation -- to something that is not a restatement of the reference by the same author.  The optional move types
are switched on as well (outer pair moves, 4-operator moves)."""
import numpy as np
import pytest

import _exact_anderson as ea

pytestmark = pytest.mark.gpu

CASES = [  # cfg (K2 = two-orbital full Kanamori, 2 x 2 blocks), bath levels per flavour, beta, extra QMC keys, scheduler
    ("C1", 3, 10.0, {}, 0),
    ("C1", 3, 10.0, {"PercentageOuterMove": 0.2, "Percentage4OperatorMove": 0.2}, 3),
    ("C2", 1, 8.0, {}, 0),
    ("K2", 2, 6.0, {}, 0),
    ("K2", 2, 6.0, {"PercentageOuterMove": 0.2, "Percentage4OperatorMove": 0.2}, 0),
    # state sampling (one outer eigenstate, WeightedOuterStateChoice): degenerate multiplets included -- statistically exact in any eigenbasis
    ("K2", 2, 6.0, {"statesampling": 1, "PercentageOuterMove": 0.2, "Percentage4OperatorMove": 0.2}, 0),
    ("C2", 1, 8.0, {"statesampling": 1, "PercentageOuterMove": 0.2}, 3),
]


@pytest.mark.parametrize("cfg,nbath,beta,extra,scheduler", CASES, ids=["C1", "C1-all-moves-autonomous", "C2", "Kanamori2", "Kanamori2-all-moves", "Kanamori2-state-sampling-all-moves", "C2-state-sampling-autonomous"])
def test_engine_against_exact_diagonalisation(cuda_engine_factory, cfg, nbath, beta, extra, scheduler):
    if cfg == "K2":
        p = ea.kanamori2_problem(nbath, beta, Nmeas=100, Ntau=1000, **extra)
    else:
        p = ea.small_bath_problem(cfg, nbath, beta, Nmeas=100, Ntau=1000, **extra)
    gex, occ_ex, gl_ex = ea.solve_cached((cfg, nbath, beta), p)      # flavour 0; all flavours are equivalent in these models
    eng = cuda_engine_factory(p, nchains=4096, nbins=32, scheduler=scheduler)
    eng.ctqmc_warmup(6000)
    eng.ctqmc_measure(1, 1)
    assert eng.scalar("error_flags") == 0
    nb = p.nbands
    gb = eng.bins("gtau").reshape(32, 2 * nb, 50, 20).mean(3).mean(1)            # [bin, tau bin], flavour average
    g, gerr = gb.mean(0), gb.std(0, ddof=1) / np.sqrt(32.0)
    ob = eng.bins("occ")
    occ, oerr = ob.mean(0), ob.std(0, ddof=1) / np.sqrt(32.0)
    lb = eng.bins("gleg").reshape(32, 2 * nb, -1).mean(1)[:, :12]                # Legendre coefficients l < 12, flavour average
    gl, glerr = lb.mean(0), lb.std(0, ddof=1) / np.sqrt(32.0)
    zg = (g - gex[0]) / np.sqrt(gerr ** 2 + (1e-4) ** 2)
    zo = ((occ - occ_ex) / np.sqrt(oerr ** 2 + (1e-4) ** 2)).ravel()
    zl = (gl - gl_ex[0]) / np.sqrt(glerr ** 2 + (1e-4) ** 2)
    z = np.concatenate([zg, zo, zl])
    print("%s + %d bath levels, beta = %g, %s: %d values, max |z| %.2f, mean z^2 %.2f; relative G error bar %.1e; <k> %.2f; sign %.4f"
          % (cfg, nbath, beta, extra or "pair moves", z.size, np.abs(z).max(), (z ** 2).mean(), (gerr / g).mean(), eng.scalar("mean_order"), float(eng.mean_sign)))
    assert np.abs(z).max() < 5.0 and (z ** 2).mean() < 2.5
    assert (gerr / g).mean() < 0.02                                  # the comparison has teeth
    if cfg == "K2":
        # reduced density matrices incl. the spin-flip / pair-hopping elements of rho2 (index convention of Meas_rho2_from_Densitymatrix,
        # CTQMC.F90:2945-3017).  The library keeps no per-group samples of rho2; its elements are linear in the same density matrix
        # as occ, so the largest occ error bar sets the scale.
        _, _, _, r1_ex, r2_ex = ea.solve_cached((cfg, nbath, beta, "rho"), p, want_rho=True)
        nf = 2 * nb
        d1 = np.abs(np.asarray(eng.rho1).reshape(nf, nf) - r1_ex).max()
        d2 = np.abs(np.asarray(eng.rho2).reshape((nf,) * 4) - r2_ex).max()
        tol = 6.0 * oerr.max() + 2e-4
        print("  rho1 max deviation %.2e, rho2 max deviation %.2e over %d non-zero elements (tolerance %.2e)" % (d1, d2, (np.abs(r2_ex) > 1e-3).sum(), tol))
        assert d1 < tol and d2 < tol
    eng.dest_ctqmc()
