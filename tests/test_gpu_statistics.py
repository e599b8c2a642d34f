"""north_star's second correctness leg on the GPU: "Final G(tau), occupations and <k> agree with the reference's own MPI CPU run
within stated Monte Carlo error bars" -- for EVERY named configuration (C1 is in tests/test_gpu_parity.py).

The CPU side is the oracle run the way the reference runs under MPI (cthyb:144, cthyb_solver.py:361-390): P independent chains with
seeds of their own, each thermalised and measured on its own, combined as DistributedSample does (mean over chains, standard error
from the chain-to-chain scatter, statistics.py:99-132, 180-209).  The GPU side is the CUDA engine through the C ABI; its error bars
come from the chain-group bins (`bins()`).  The two sides share no random numbers (different Philox keys).
Stated criterion, per quantity: z = (gpu - cpu) / sqrt(err_gpu^2 + err_cpu^2 + floor^2);  max |z| < 5 and mean z^2 < 2.5
over all compared values (a few hundred per configuration), floor = 2e-4 absolute (protects bins whose scatter estimate is itself
noisy because they hold almost no weight)."""
import os
import threading

import numpy as np
import pytest

from oracle.oracle import OracleChain
from w2dynamics_b200 import problems

pytestmark = pytest.mark.gpu

NT = 25                      # coarse tau bins compared


def oracle_population(p, nchains, warm, nmeas, fast):
    nthreads = min(nchains, os.cpu_count() or 1)
    chains = [OracleChain(p, rng="philox", seed=777, chain=50000 + i, fast_trace=fast) for i in range(nchains)]

    def work(t):
        for c in chains[t::nthreads]:
            c.warmup(warm)
            c.measure(nmeas)
    ths = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert all(c.error == 0 for c in chains)
    return [c.results() for c in chains]


def coarse(g):
    s = g.shape
    return g.reshape(s[:-1] + (NT, s[-1] // NT)).mean(-1)


def order(hist):
    return (hist * np.arange(hist.shape[-1])).sum(-1)


def stats(samples):
    a = np.array(samples)
    return a.mean(0), a.std(0, ddof=1) / np.sqrt(a.shape[0])


CASES = {
    #      gpu chains, gpu warm, gpu Nmeas | cpu chains, cpu warm, cpu Nmeas, oracle trace mode
    "C2": (2048, 20000, 100, 32, 30000, 2000, 0),
    "C3": (2048, 12000, 100, 32, 15000, 1500, 2),
    "C4": (2048, 12000, 100, 48, 12000, 2500, 2),
    "C5": (1024, 80000, 60, 32, 80000, 500, 0),
}


@pytest.mark.parametrize("cfg", sorted(CASES))
def test_statistical_agreement_with_cpu_population(cuda_engine_factory, cfg):
    gch, gwarm, gmeas, cch, cwarm, cmeas, fast = CASES[cfg]
    p = problems.make_problem(cfg, Nmeas=gmeas, Ntau=500, NLegMax=40)
    _compare(cuda_engine_factory, cfg, p, p, CASES[cfg])


@pytest.mark.parametrize("cfg,beta", [("C3", 3.0), ("C2", 4.0)])
def test_outer_moves_leave_the_distribution_invariant(cuda_engine_factory, cfg, beta):
    """Outer pair moves (PercentageOuterMove) only add a way to move through configuration space: the GPU population
    sampled WITH 30 % outer moves must agree with a CPU oracle population sampled WITHOUT them, inside the same error bars (high
    temperature, where a tenth of these moves is accepted)."""
    case = (2048, 6000, 100, 32, 8000, 3000, 0)
    p_gpu = problems.make_problem(cfg, beta=beta, Nmeas=case[2], Ntau=500, NLegMax=40, PercentageOuterMove=0.3)
    p_cpu = problems.make_problem(cfg, beta=beta, Nmeas=case[2], Ntau=500, NLegMax=40)
    _compare(cuda_engine_factory, "%s beta=%g with outer moves" % (cfg, beta), p_gpu, p_cpu, case)


@pytest.mark.parametrize("cfg,beta", [("C3", None), ("C3", 3.0), ("C2", 4.0)])
def test_four_operator_moves_leave_the_distribution_invariant(cuda_engine_factory, cfg, beta):
    """4-operator moves (Percentage4OperatorMove, the move half of SURVEY.md 8 f3): the GPU population sampled WITH 30 % of the pair
    moves replaced by two-pair moves must agree with a CPU oracle population sampled WITHOUT them (C3 as named: 4 % of them are
    accepted; high temperature: 10 %)."""
    case = (2048, 12000, 100, 32, 15000, 1500, 2) if beta is None else (2048, 6000, 100, 32, 8000, 3000, 0)
    p_gpu = problems.make_problem(cfg, beta=beta, Nmeas=case[2], Ntau=500, NLegMax=40, Percentage4OperatorMove=0.3)
    p_cpu = problems.make_problem(cfg, beta=beta, Nmeas=case[2], Ntau=500, NLegMax=40)
    _compare(cuda_engine_factory, "%s beta=%s with 4-operator moves" % (cfg, beta), p_gpu, p_cpu, case)


def _compare(cuda_engine_factory, cfg, p, p_cpu, case):
    gch, gwarm, gmeas, cch, cwarm, cmeas, fast = case
    eng = cuda_engine_factory(p, nchains=gch, nbins=32)
    eng.ctqmc_warmup(gwarm)
    eng.ctqmc_measure(1, 1)
    assert eng.scalar("error_flags") == 0
    nb = p.nbands
    # ---- GPU: population result + error bars over the 32 chain-group bins
    gb, ob = eng.bins("gtau"), eng.bins("occ")
    g_gpu, g_gerr = coarse(eng.gtau), coarse(gb).std(0, ddof=1) / np.sqrt(gb.shape[0])
    occ_gpu, occ_gerr = eng.occ, ob.std(0, ddof=1) / np.sqrt(ob.shape[0])
    k_gpu, sign_gpu = order(eng.histo), float(eng.mean_sign)
    sb = eng.bins("sign")
    sign_gerr = sb.std(ddof=1) / np.sqrt(sb.size)
    eng.dest_ctqmc()
    # ---- CPU: independent oracle chains, one sample each (DistributedSample semantics)
    res = oracle_population(p_cpu, cch, cwarm, cmeas, fast)
    g_cpu, g_cerr = stats([coarse(r["gtau"]) for r in res])
    occ_cpu, occ_cerr = stats([r["occ"] for r in res])
    k_cpu, k_cerr = stats([order(r["hist"]) for r in res])
    sign_cpu, sign_cerr = stats([r["sign"] for r in res])
    floor = 2e-4
    zs = [(g_gpu - g_cpu) / np.sqrt(g_gerr ** 2 + g_cerr ** 2 + floor ** 2)]
    d = np.arange(nb)
    for s1 in range(2):                                                        # occupations and same-orbital double occupancy
        zs.append((occ_gpu[d, s1, d, s1] - occ_cpu[d, s1, d, s1]) / np.sqrt(occ_gerr[d, s1, d, s1] ** 2 + occ_cerr[d, s1, d, s1] ** 2 + floor ** 2))
    zs.append((occ_gpu[d, 0, d, 1] - occ_cpu[d, 0, d, 1]) / np.sqrt(occ_gerr[d, 0, d, 1] ** 2 + occ_cerr[d, 0, d, 1] ** 2 + floor ** 2))
    if nb > 1:                                                                 # inter-orbital density correlations
        zs.append(((occ_gpu[0, :, 1, :] - occ_cpu[0, :, 1, :]) / np.sqrt(occ_gerr[0, :, 1, :] ** 2 + occ_cerr[0, :, 1, :] ** 2 + floor ** 2)).ravel())
    # <k_f>: the GPU histogram has no bins -> its error from the CPU scatter scaled by the sample-size ratio (same estimator, same chain)
    k_gerr = k_cerr * np.sqrt(cch * cmeas / float(gch * gmeas))
    zs.append(((k_gpu - k_cpu) / np.sqrt(k_gerr ** 2 + k_cerr ** 2 + 1e-6)).ravel())
    zs.append(np.array([(sign_gpu - sign_cpu) / np.sqrt(sign_gerr ** 2 + sign_cerr ** 2 + 1e-8)]))
    z = np.concatenate([np.asarray(x).ravel() for x in zs])
    print("%s: %d values compared, max |z| %.2f, mean z^2 %.2f; <k> gpu %s cpu %s; sign gpu %.4f cpu %.4f +- %.4f; G(tau) errors gpu %.1e cpu %.1e"
          % (cfg, z.size, np.abs(z).max(), (z ** 2).mean(), np.round(k_gpu.ravel()[:2], 3), np.round(k_cpu.ravel()[:2], 3), sign_gpu, sign_cpu, sign_cerr,
             g_gerr.mean(), g_cerr.mean()))
    assert np.all(np.isfinite(z))
    assert np.abs(z).max() < 5.0 and (z ** 2).mean() < 2.5, (np.abs(z).max(), (z ** 2).mean())
    # the comparison has teeth: the CPU error bars resolve G(tau) to a few per cent
    assert g_cerr.mean() < 0.08 * np.abs(g_cpu).mean() + 1e-3
