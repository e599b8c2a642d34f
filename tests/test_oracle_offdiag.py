"""ORACLE-ONLY groundwork for `offdiag = 1` (the open half of SURVEY.md 8 f3; the CUDA engine still refuses the key): one matrix over
all flavours in the order of `get_FullHybr_offdiag` (Trace.F90:2690-2770), pair moves whose creator draws its own flavour
(`pair_OperAdd` with force_diagonal = .false., Trace.F90:3657-3698; `(2 NBands)^2 / NPairs`, CTQMC.F90:3517-3519), the flavour-blind pair
count (`is_rempair_offdiag`, Trace.F90:2018-2030), `recalc_sign` (Trace.F90:3500-3546), 4-operator moves, global moves (CTQMC.F90:4491-4494), tau shifts, `MeasGtauRem_full`
(CTQMC.F90:2475-2584) with its normalisation (8752-8753); determinants from scratch (the reference's own `b_full_offdiag` branch,
CTQMC.F90:3482-3487).  Pinned to EXACT results of a two-orbital impurity whose bath levels couple to both orbitals: the full G_ab(tau)
matrix -- its orbital-off-diagonal elements change sign along tau -- and all <n_i n_j>."""
import threading

import numpy as np
import pytest

import _exact_anderson as ea


def _run(p, nchains=32, warm=4000, meas=8000):
    from oracle.oracle import OracleChain
    chains = [OracleChain(p, rng="philox", seed=4242, chain=7000 + i) for i in range(nchains)]

    def work(t):
        for c in chains[t::8]:
            c.warmup(warm)
            c.measure(meas)
    ths = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert all(c.error == 0 for c in chains)
    return [c.results() for c in chains]


def _compare(res, pairs, gex, occ_ex, label):
    n = len(res)
    gf = np.array([r["gtau-full"].reshape(4, 4, 50, 20).mean(3) for r in res])
    zs = []
    for (a, b), ge in zip(pairs, gex):
        g, e = gf[:, a, b].mean(0), gf[:, a, b].std(0, ddof=1) / np.sqrt(n)
        zs.append((g - ge) / np.sqrt(e ** 2 + 1e-8))
    z = np.concatenate(zs)
    occ = np.array([r["occ"] for r in res])
    zo = ((occ.mean(0) - occ_ex) / np.sqrt(occ.var(0, ddof=1) / n + 1e-8)).ravel()
    zoff = np.concatenate([zs[i] for i, (a, b) in enumerate(pairs) if a != b])
    print("%s: G_ab %d values max |z| %.2f mean z^2 %.2f (off-diagonal elements alone: %.2f / %.2f); occupations max |z| %.2f; sign %.3f"
          % (label, z.size, np.abs(z).max(), (z ** 2).mean(), np.abs(zoff).max(), (zoff ** 2).mean(), np.abs(zo).max(), np.mean([r["sign"] for r in res])))
    assert np.abs(z).max() < 5.0 and (z ** 2).mean() < 2.0 and np.abs(zo).max() < 4.5
    assert np.abs(gex[[i for i, (a, b) in enumerate(pairs) if a != b]]).max() > 0.005          # the off-diagonal elements are not noise


PAIRS = [(0, 0), (2, 2), (1, 1), (0, 2), (2, 0), (1, 3)]


def test_offdiag_noninteracting_exact(built):
    """U = 0: the exact G_ab(tau) follows from the one-particle Hamiltonian of impurity + bath (checked against the Fock-space solver here)."""
    p = ea.offdiag_two_orbital_problem(5.0, NCorr=20, PercentageTauShiftMove=0.02, PercentageGlobalMove=0.02)
    p.u_matrix[...] = 0.0
    for b in range(2):
        for s in range(2):
            p.muimp[b, s, b, s] = 0.3 * (b - 0.5)
    gex, occ_ex = ea.solve_general(p, PAIRS)
    h = np.zeros((5, 5))
    h[0, 0], h[1, 1] = p.muimp[0, 0, 0, 0], p.muimp[1, 0, 1, 0]
    for q in range(3):
        h[2 + q, 2 + q] = p.bath_eps[q]
        h[0, 2 + q] = h[2 + q, 0] = p.bath_v[0, q]
        h[1, 2 + q] = h[2 + q, 1] = p.bath_v[2, q]
    w, u = np.linalg.eigh(h)
    tau = (np.arange(50) + 0.5) * p.beta / 50
    g01 = (u[0][None, :] * u[1][None, :] * np.exp(-tau[:, None] * w[None, :]) / (1 + np.exp(-p.beta * w[None, :]))).sum(1)
    assert np.abs(gex[3] - g01).max() < 3e-4 and g01.max() > 0.03 and g01.min() < -0.02       # bin average vs midpoint value
    _compare(_run(p), PAIRS, gex, occ_ex, "offdiag oracle, U = 0")


def test_offdiag_interacting_exact(built):
    """Density-density interaction (pair moves are ergodic: every flavour's particle number is conserved by H_loc; see the note in
    tests/_exact_anderson.py on Kanamori terms)."""
    p = ea.offdiag_two_orbital_problem(5.0, NCorr=20, PercentageTauShiftMove=0.02, PercentageGlobalMove=0.02)
    gex, occ_ex = ea.solve_general(p, PAIRS)
    _compare(_run(p, nchains=48, meas=12000), PAIRS, gex, occ_ex, "offdiag oracle, density-density U = 4")


def test_offdiag_kanamori_with_four_operator_moves_exact(built):
    """Full Kanamori interaction: H_loc does not conserve the particle number per flavour, flavour-unbalanced configurations carry weight
    through F_01 and only the 4-operator moves (StepAdd4 / StepRem4 with `offdiag`: four independent flavours, (2 NBands)^4, CTQMC.F90:3641,
    3823, 4353) reach them -- with them the whole G_ab(tau) matrix is exact, at a mean sign of -0.97 (recalc_sign at work); without them
    G_01 comes out 3-5 x too large (note in tests/_exact_anderson.py)."""
    p = ea.offdiag_two_orbital_problem(5.0, kanamori=True, NCorr=20, PercentageTauShiftMove=0.02, PercentageGlobalMove=0.02, Percentage4OperatorMove=0.3)
    gex, occ_ex = ea.solve_general(p, PAIRS)
    res = _run(p, nchains=48, warm=6000, meas=16000)
    sign = np.mean([r["sign"] for r in res])
    assert -0.995 < sign < -0.9                                    # a genuine, mild sign problem
    n = len(res)
    gf = np.array([r["gtau-full"].reshape(4, 4, 50, 20).mean(3) for r in res])
    zs = []
    for (a, b), ge in zip(PAIRS, gex):
        g, e = gf[:, a, b].mean(0), gf[:, a, b].std(0, ddof=1) / np.sqrt(n)
        zs.append((g - ge) / np.sqrt(e ** 2 + 1e-8))
    z = np.concatenate(zs)
    occ = np.array([r["occ"] for r in res])
    zo = ((occ.mean(0) - occ_ex) / np.sqrt(occ.var(0, ddof=1) / n + 1e-8)).ravel()
    print("offdiag oracle, Kanamori + 4-operator moves: G_ab %d values max |z| %.2f mean z^2 %.2f; occupations max |z| %.2f; sign %.3f"
          % (z.size, np.abs(z).max(), (z ** 2).mean(), np.abs(zo).max(), sign))
    assert np.abs(z).max() < 5.0 and (z ** 2).mean() < 2.0 and np.abs(zo).max() < 4.5


def test_offdiag_rank1_ratio_equals_from_scratch(built):
    """The determinant ratio of every pair proposal by the rank-1 form on the stored single inverse (`get_LogDetRatPairAdd_full /
    Rem_full`, Trace.F90:2835-2997: S = F_new - R M Q over all flavours resp. the inverse's element, sign (-1)^(row + column) in the
    flavour-major order) equals the from-scratch determinant ratio the oracle walks with -- the formula the engine's flavour-agnostic
    `det_ratio_add` has to reproduce for `offdiag = 1`."""
    from oracle.oracle import OracleChain
    p = ea.offdiag_two_orbital_problem(5.0, kanamori=True, NCorr=20, PercentageTauShiftMove=0.02, PercentageGlobalMove=0.02, Percentage4OperatorMove=0.2)
    o = OracleChain(p, rng="philox", seed=3, chain=1)
    log = o.steps(20000, log=True)
    pm = log[((log[:, 0] == 1) | (log[:, 0] == 2)) & (log[:, 1] != 2) & np.isfinite(log[:, 7])]
    assert pm.shape[0] > 2000 and (pm[:, 1] == 1).sum() > 300 and (pm[:, 8] < 0).sum() > 500       # many proposals, both signs
    assert np.abs(pm[:, 7] - pm[:, 18]).max() < 1e-9 and np.array_equal(pm[:, 8], pm[:, 19])
    assert o.error == 0


def test_offdiag_oracle_refuses_unrestated_moves(built):
    from oracle.oracle import OracleChain
    p = ea.offdiag_two_orbital_problem(5.0, PercentageOuterMove=0.1)
    with pytest.raises(ValueError, match="offdiag"):
        OracleChain(p, seed=1)
