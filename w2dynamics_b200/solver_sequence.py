"""Host-side mirror of the reference's solver plugin for the hot path: the call sequence `CtHybSolver` issues against the
`mctqmc` module object, call for call (the reference's plugin itself needs configobj / mpi4py / h5py and the w2dyn package; where
those exist, redirect its one import as INTEGRATION.md shows and use it unchanged):

    CtHybSolver.__init__           cthyb_solver.py:99-114    init_rnd(seed), fancy_prog
    CtHybSolver.set_problem        cthyb_solver.py:223-260   parameter string (764-774), init_paras, fourpnt, array conventions
    CtHybSolver.solve              cthyb_solver.py:316-400   simid, init_solver, [set_mc_config], set_taudiffmax, ctqmc_warmup,
                                                             ctqmc_measure(1,1), get_mc_config, aborted, results, dest_*
    calibrate_taudiff_max          cthyb_solver.py:899-944   two ctqmc_calibrate phases + exponential fit
    get_resultset                  cthyb_solver.py:946-1032  the module arrays that are read, in the reference's order

`tests/test_reference_plugin.py::test_restated_sequence_equals_plugin` runs the REAL plugin and this mirror against a recording
proxy and asserts that both issue the same calls with the same arguments; `bench.py`'s end-to-end figure and the `-m gpu` boundary
tests drive the CUDA library with it.
"""
import time

import numpy as np

RESULT_READS = [  # get_resultset, cthyb_solver.py:950-980 (+ 1016-1026), in order
    ("gtau", "gtau"), ("gtau-full", "gtau_full"), ("rhist", "rhisto"), ("lhist", "lhisto"), ("hist", "histo"),
    ("ssts-states", "statessuperstates"), ("energies-eigenstates", "eigenstatesenergies"), ("occbasis-mapping", "occbasismapping"),
    ("hist-sst", "outersuperstatehisto"), ("hist-state", "outerstatehisto"), ("sign-sst", "signhistosuperstates"),
    ("sign-state", "signhistostates"), ("contrib-sst", "tracecontribsuperstates"), ("contrib-state", "tracecontribstates"),
    ("occ", "occ"), ("gleg", "gleg"), ("gleg-full", "gleg_full"), ("accept-ins", "accadd"), ("accept-ins4", "accadd4"),
    ("accept-rem", "accrem"), ("accept-rem4", "accrem4"), ("accept-pair-tau", "accpair"), ("accept-glob", "accglob"),
    ("accept-shift", "accshift"), ("accept-flavourchange", "accflavc"), ("sign", "mean_sign"), ("time-warmup", "time_warmup"),
    ("time-simulation", "time_sim"), ("time-sampling", "time_sim_steps"),
]


class MCConfig:             # CtHybConfig, cthyb_solver.py:24-86
    def __init__(self, taus, orbs, spins, cas, hashybs, outer_sst, outer_state):
        self.taus, self.orbs, self.spins, self.cas, self.hashybs = taus, orbs, spins, cas, hashybs
        self.outer_sst, self.outer_state = outer_sst, outer_state

    @classmethod
    def get_from_ctqmc(cls, solver):                 # 46-56
        noper, outer_sst, outer_state = solver.get_mc_config_scalars()
        if noper > 0:
            return cls(*solver.get_mc_config(noper))
        e = np.array((), dtype=int)
        return cls(np.array((), dtype=np.double), e, e, e, e, outer_sst, outer_state)

    def set_to_ctqmc(self, solver):                  # 58-65
        if self.taus.size > 0:
            solver.set_mc_config(self.taus.size, self.taus, self.orbs, self.spins, self.cas, self.hashybs, self.outer_sst, self.outer_state)
        else:
            solver.set_empty_mc_config(self.outer_sst, self.outer_state)


class PluginSequence:
    """What CtHybSolver does to `ctqmc`, for a diagonal-hybridisation problem on one rank."""

    def __init__(self, ctqmc, qmc_config, seed, interactive=False):       # cthyb_solver.py:99-114
        self.ctqmc, self.qmc = ctqmc, dict(qmc_config)
        self.param_string = "".join("%s = %s #\n" % (k, v) for k, v in sorted(self.qmc.items()))     # config_to_pstring 764-767
        ctqmc.init_rnd(seed)
        ctqmc.fancy_prog = interactive

    def set_problem(self, hamiltonian, beta, norbitals, quantum_numbers, u_matrix, ftau_tau_first, muimp_lattice, screening_tau_first,
                    symmetry_moves=()):
        """ftau_tau_first / screening_tau_first: (Nftau, Nb, 2, Nb, 2) as ImpurityProblem holds them; muimp_lattice: lattice
        convention (+mu)."""
        problem_params = {"Hamiltonian": hamiltonian, "beta": beta, "Nd": norbitals, "ParaMag": 0, "QuantumNumbers": quantum_numbers,
                          "Phonon": 0, "g_phonon": "", "omega0": "", "Screening": 0, "Uw": 0}
        s = self.param_string + "".join("%s = %s #\n" % (k, v) for k, v in sorted(problem_params.items()))
        s += "NSymMove = %d #\n" % len(symmetry_moves)                                                # symm_to_pstring 769-774
        s += "".join("SymMove%02d = %s #\n" % (i + 1, " ".join(map(str, np.asarray(v) + 1))) for i, v in enumerate(symmetry_moves))
        self.ctqmc.init_paras(s)
        self.ctqmc.fourpnt = 0
        nb = norbitals
        ft, mu = np.array(ftau_tau_first), np.array(muimp_lattice)
        if self.qmc["offdiag"] == 0:                 # promote_diagonal(extract_diagonal(.)) 244-246
            d = np.zeros_like(ft)
            m = np.zeros_like(mu)
            for b in range(nb):
                for sp in range(2):
                    d[:, b, sp, b, sp] = ft[:, b, sp, b, sp]
                    m[b, sp, b, sp] = mu[b, sp, b, sp]
            ft, mu = d, m
        self.ftau = ft.transpose(1, 2, 3, 4, 0)      # tau axis last 252
        self.muimp = np.real(-mu)                    # 255-262
        self.umatrix = np.real(np.asarray(u_matrix).reshape((2 * nb,) * 4))
        self.screening = np.asarray(screening_tau_first).transpose(1, 2, 3, 4, 0).real                # 278-279
        self.beta = beta

    def calibrate_taudiff_max(self):                 # 899-944
        import scipy.optimize as opt
        c = self.ctqmc
        phase1pairnum, phase2taugrid, phase2pairnum, nprogress = 1000, 100, 1000, 100000
        c.init_counters()
        c.ctqmc_calibrate(True, phase1pairnum, phase2taugrid, phase2pairnum, nprogress)
        acctaulist = c.accpairtau
        acctaulist.sort()
        taudiffmax = float(acctaulist[int(0.995 * phase1pairnum)])
        c.set_taudiffmax(taudiffmax)
        c.init_counters()
        c.ctqmc_calibrate(False, phase1pairnum, phase2taugrid, phase2pairnum, nprogress)
        accept_pair = np.asarray(c.accpair, dtype=float)
        loc = 0.5 * taudiffmax / phase2taugrid                                                        # tf.tau_bins(taudiffmax, grid, "centre")
        taus = np.linspace(loc, taudiffmax + loc, phase2taugrid, endpoint=False)                      # transform.py:72-86
        est = np.empty(accept_pair.shape[:2])
        for b in range(accept_pair.shape[0]):
            for s in range(accept_pair.shape[1]):
                fit = opt.leastsq(lambda p, t, a: a - p[0] * np.exp(p[1] * t), (accept_pair[b, s, 0], -1), (taus, accept_pair[b, s]))
                est[b, s] = np.log(1 - 0.99) / fit[0][1]
        return np.amax(est)

    def solve(self, iter_no, mc_config_inout=None, mpi_rank=0):                                      # 316-400
        c, q = self.ctqmc, self.qmc
        c.simid = mpi_rank
        t0 = time.perf_counter()
        c.init_solver(self.umatrix, self.ftau, self.muimp, self.screening, iter_no + 1)
        firstrun = True
        if q["ReuseMCConfig"] != 0 and type(mc_config_inout) is list and len(mc_config_inout) > 0:
            firstrun = False
            mc_config_inout.pop(0).set_to_ctqmc(c)
        if q["segment"] == 0:
            c.set_taudiffmax(q["TaudiffMax"] if q["TaudiffMax"] > 0.0 else self.calibrate_taudiff_max())
        c.ctqmc_warmup(q["Nwarmups"] if firstrun or q["Nwarmups2Plus"] < 0 else q["Nwarmups2Plus"])
        c.ctqmc_measure(1, 1)
        dt = time.perf_counter() - t0
        if type(mc_config_inout) is list:
            mc_config_inout.append(MCConfig.get_from_ctqmc(c))
        self.abort = False
        if c.aborted == 2:  # no data, kill run                                                      # 347-352 (two reads)
            raise KeyboardInterrupt("CTQMC was interrupted by signal")
        elif c.aborted == 1:
            self.abort = True
        result = self.get_resultset()
        result["time-qmc"] = dt
        c.dest_ctqmc()
        c.dest_paras()
        return result

    def get_resultset(self):                         # 946-1032
        c, q = self.ctqmc, self.qmc
        norbitals = c.gtau.shape[0]
        result = {}
        for key, attr in RESULT_READS:
            result[key] = getattr(c, attr)
        if q.get("MeasGiw", 0) != 0 and q["offdiag"] == 0:                   # 989-990
            result["giw-meas"] = c.giw
        if q["MeasDensityMatrix"] != 0 and q["segment"] == 0:
            result["densitymatrix"] = c.densitymatrix
            result["rho2"] = c.rho2.reshape((norbitals, 2) * 4)
            result["rho1"] = c.rho1.reshape((norbitals, 2) * 2)
        if q["segment"] == 0:
            result["hist"] = c.histo
        if q["AccumGtau"] == 2:
            result["gtau-blocks"] = c.gtau_blocks
        return dict((k, np.array(result[k]).copy()) for k in result)


def qmc_defaults():
    """[QMC] defaults of the reference's configspec that the sequence above consults (configspec lines in SURVEY.md 5); the full
    section is parsed from the reference where it is present (tests/test_reference_plugin.py::_qmc_defaults)."""
    from .problems import QMC_DEFAULTS
    q = dict(QMC_DEFAULTS)
    for k, v in (("ReuseMCConfig", 0), ("segment", 0), ("offdiag", 0), ("Nwarmups2Plus", -1), ("MeasDensityMatrix", 1), ("AccumGtau", 1)):
        q.setdefault(k, v)
    return q
