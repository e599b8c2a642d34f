"""Synthetic impurity problems C1..C5 for the CT-HYB hot path.

Everything here is deterministic input generation; no random numbers enter
the *inputs*.  The conventions follow the reference's single-shot ``cthyb``
route with ``General.DOS = EDcheck`` (reference ``cthyb:170-198``):

* hole propagator ``F(tau) = sum_p V_p^2 / (exp(-tau e_p) + exp((beta-tau) e_p))``
  on ``tau_i = i*beta/(Nftau-1)`` (``w2dyn/auxiliaries/transform.py:114-146``),
* U-matrices in the reference's ``(band, spin)^4`` layout reshaped to
  ``(nf, nf, nf, nf)`` with flavour ``2*band + spin``
  (``w2dyn/dmft/interaction.py:171-273, 275-425``),
* ``muimp`` handed to the engine already negated (``cthyb_solver.py:255``).

The U-matrix builders are our own implementations; ``tests/test_oracle_pins.py``
pins them against vectors generated from the reference's ``interaction.py``
(``tests/golden/umatrix.npz``, generator ``tools/gen_golden.py``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np


# ----------------------------------------------------------------------------
# U matrices
# ----------------------------------------------------------------------------
def u_density(nb: int, u: float, v: float, j: float) -> np.ndarray:
    """Density-density interaction, layout U[b1,s1,b2,s2,b3,s3,b4,s4]
    (reference interaction.py:182-189)."""
    um = np.zeros((nb, 2, nb, 2, nb, 2, nb, 2))
    for a in range(nb):
        for s in range(2):
            um[a, s, a, 1 - s, a, s, a, 1 - s] = u
        for b in range(nb):
            if a == b:
                continue
            for s in range(2):
                for t in range(2):
                    um[a, s, b, t, a, s, b, t] = v
                um[a, s, b, s, b, s, a, s] = j
    return um


def u_kanamori(nb: int, u: float, v: float, j: float) -> np.ndarray:
    """Kanamori = density-density + pair hopping + spin flip
    (reference interaction.py:232-241)."""
    um = u_density(nb, u, v, j)
    for a in range(nb):
        for b in range(nb):
            if a == b:
                continue
            for s in range(2):
                um[a, s, a, 1 - s, b, s, b, 1 - s] = j  # pair hopping
                um[a, s, b, 1 - s, b, s, a, 1 - s] = j  # spin flip
    return um


def _clebsch_gordan(j1, m1, j2, m2, J, M) -> float:
    """<j1 m1 j2 m2 | J M> via Racah's closed formula (integer momenta)."""
    if m1 + m2 != M or abs(m1) > j1 or abs(m2) > j2 or abs(M) > J:
        return 0.0
    if J > j1 + j2 or J < abs(j1 - j2):
        return 0.0
    f = math.factorial
    pref = (2 * J + 1) * f(J + j1 - j2) * f(J - j1 + j2) * f(j1 + j2 - J) / f(j1 + j2 + J + 1)
    pref *= f(J + M) * f(J - M) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2)
    tot = 0.0
    for k in range(0, j1 + j2 - J + 1):
        d = [k, j1 + j2 - J - k, j1 - m1 - k, j2 + m2 - k, J - j2 + m1 + k, J - j1 - m2 + k]
        if min(d) < 0:
            continue
        den = 1.0
        for x in d:
            den *= f(x)
        tot += (-1.0) ** k / den
    return math.sqrt(pref) * tot


def u_coulomb(nb: int, f0: float, f2: float, f4: float, f6: float = 0.0) -> np.ndarray:
    """Full Slater-Coulomb interaction in cubic harmonics
    (restates the convention of reference interaction.py:328-419:
    a_k(m1,m2,m3,m4) = sum_q <l m3 k q|l m1><l m2 k q|l m4> <l0 k0|l0>^2,
    cubic order for d: xy, yz, z^2, xz, x^2-y^2)."""
    if nb not in (1, 3, 5):
        raise NotImplementedError("Coulomb builder: nb in {1,3,5}")
    l = (nb - 1) // 2
    ms = range(-l, l + 1)
    sph = np.zeros((nb,) * 4)
    slater = {0: f0, 2: f2, 4: f4, 6: f6}
    for k, fk in slater.items():
        if fk == 0.0 or k > 2 * l:
            continue
        red = _clebsch_gordan(l, 0, k, 0, l, 0) ** 2
        for m1 in ms:
            for m2 in ms:
                for m3 in ms:
                    for m4 in ms:
                        a = 0.0
                        for q in range(-k, k + 1):
                            a += (_clebsch_gordan(l, m3, k, q, l, m1)
                                  * _clebsch_gordan(l, m2, k, q, l, m4))
                        sph[m1 + l, m2 + l, m3 + l, m4 + l] += fk * a * red
    rot = np.zeros((nb, nb), dtype=complex)
    s2 = 1.0 / math.sqrt(2.0)
    if nb == 5:
        rot[0, 0] = 1j * s2; rot[0, 4] = -1j * s2
        rot[1, 1] = 1j * s2; rot[1, 3] = 1j * s2
        rot[2, 2] = 1.0
        rot[3, 1] = s2; rot[3, 3] = -s2
        rot[4, 0] = s2; rot[4, 4] = s2
    elif nb == 3:
        rot[0, 0] = 1j * s2; rot[0, 2] = 1j * s2
        rot[1, 1] = 1.0
        rot[2, 0] = s2; rot[2, 2] = -s2
    else:
        rot[0, 0] = 1.0
    cub = np.einsum("ai,bj,ijkl,ck,dl->abcd", rot.conj(), rot.conj(), sph, rot, rot)
    um = np.zeros((nb, 2, nb, 2, nb, 2, nb, 2))
    for s in range(2):
        for t in range(2):
            um[:, s, :, t, :, s, :, t] = cub.real
    return um, cub


def udens_from_ufull(cub: np.ndarray) -> np.ndarray:
    """U_{i s, j t} = U_ijij - U_ijji delta_st (reference interaction.py:57-74)."""
    nb = cub.shape[0]
    ud = np.zeros((nb, 2, nb, 2))
    for i in range(nb):
        for j in range(nb):
            for s in range(2):
                for t in range(2):
                    ud[i, s, j, t] = cub[i, j, i, j].real - (cub[i, j, j, i].real if s == t else 0.0)
    return ud


# ----------------------------------------------------------------------------
# bath and hybridisation
# ----------------------------------------------------------------------------
def bethe_bath(n_bath: int, half_bandwidth: float = 2.0):
    """Gauss-Chebyshev (2nd kind) discretisation of the Bethe semicircle
    (SURVEY.md §8(d)): sum V^2 = D^2/4."""
    p = np.arange(1, n_bath + 1)
    x = p * math.pi / (n_bath + 1)
    eps = half_bandwidth * np.cos(x)
    v2 = (half_bandwidth ** 2 / 4.0) * (2.0 / (n_bath + 1)) * np.sin(x) ** 2
    return eps, np.sqrt(v2)


def hybr_tau(eps: np.ndarray, v: np.ndarray, beta: float, nftau: int) -> np.ndarray:
    """F(tau_i) on the closed grid tau_i = i beta/(nftau-1)
    (restates transform.hybr_from_sites, transform.py:140-143)."""
    tau = np.linspace(0.0, beta, nftau)
    den = np.exp(-tau[:, None] * eps[None, :]) + np.exp((beta - tau[:, None]) * eps[None, :])
    return ((v ** 2)[None, :] / den).sum(-1)


# ----------------------------------------------------------------------------
# problem definitions
# ----------------------------------------------------------------------------
QMC_DEFAULTS = {
    # hot-path keys with the reference's configspec defaults
    "Nwarmups": 100000, "Nmeas": 10000, "NCorr": 100, "NSeed": 43890,
    "MaxHisto": 500, "TaudiffMax": -1.0, "MeasDensityMatrix": 1,
    "PercentageGlobalMove": 0.005, "PercentageTauShiftMove": 0.005,
    "PercentageOuterMove": 0.0, "Percentage4OperatorMove": 0.0,
    "PercentageWormInsert": 0.0, "PercentageWormReplace": 0.0,
    "Ntau": 1000, "Nftau": 5000, "Niw": 2000, "NLegMax": 100, "NLegOrder": 100,
    "AccumGtau": 1, "MeasGtauDetRat": 0, "statesampling": 0,
    "EnableOuterTruncation": "", "truncation": 1, "EPSBLOCK": 1e-13,
    "EPSEVEC": 1e-15, "EPSTRACEVEC": 1e-15, "EPSSANDWICH": 1e-15, "EPSLANC": 1e-15,
    "offdiag": 0, "full_offdiag": 0, "segment": 0, "Eigenbasis": 1,
    "flavourchange_moves": 0, "FixedPreallocatedVecs": 0, "MeasSsquare": 0,
    "PrintDensityMatrixBasis": 0, "NLookup_nfft": 100000,
    "N4leg": 1, "N4tau": 1, "N4iwb": 0, "N4iwf": 1, "N2iwb": 0, "N3iwb": 0, "N3iwf": 1,
    "Gtau_mean_step": 0, "Gtau_mid_step": 0, "sign_step": 0,
    "MeasGiw": 0, "MeasG2iw": 0, "MeasSusz": 0, "MeasExpResDensityMatrix": 0,
    "FourPnt": 0, "MeasurementTiming": -1, "ReuseMCConfig": 0,
}


@dataclass
class Problem:
    name: str
    nbands: int
    beta: float
    hamiltonian: str
    quantum_numbers: str
    u_matrix: np.ndarray          # (nf,nf,nf,nf), flavour = 2*band+spin
    muimp: np.ndarray             # (nb,2,nb,2), already negated (engine convention)
    ftau: np.ndarray              # (nb,2,nb,2,Nftau)
    qmc: dict = field(default_factory=dict)
    eps_bath: np.ndarray | None = None
    v_bath: np.ndarray | None = None
    mu: float = 0.0
    symmetry_moves: list = field(default_factory=list)   # 0-based flavour permutations (index band + nbands * spin), as ImpurityProblem.symmetry_moves

    @property
    def nflavours(self):
        return 2 * self.nbands

    def pstring(self, **overrides) -> str:
        """Parameter string exactly as ``CtHybSolver.config_to_pstring`` builds it
        (cthyb_solver.py:764-774, 226-240): sorted [QMC] keys, then problem keys,
        then NSymMove."""
        qmc = dict(QMC_DEFAULTS)
        qmc.update(self.qmc)
        qmc.update(overrides)
        s = "".join("%s = %s #\n" % (k, v) for k, v in sorted(qmc.items()))
        prob = {
            "Hamiltonian": self.hamiltonian, "beta": self.beta, "Nd": self.nbands,
            "ParaMag": 0, "QuantumNumbers": self.quantum_numbers, "Phonon": 0,
            "g_phonon": "", "omega0": "", "Screening": 0, "Uw": 0,
        }
        s += "".join("%s = %s #\n" % (k, v) for k, v in sorted(prob.items()))
        s += "NSymMove = %d #\n" % len(self.symmetry_moves)                       # symm_to_pstring, cthyb_solver.py:769-774
        s += "".join("SymMove%02d = %s #\n" % (i + 1, " ".join(str(int(v) + 1) for v in vals)) for i, vals in enumerate(self.symmetry_moves))
        return s


def _assemble(name, nb, beta, ham, qns, um8, mu, n_bath, qmc, nftau=5000):
    eps, v = bethe_bath(n_bath)
    f = hybr_tau(eps, v, beta, nftau)
    ftau = np.zeros((nb, 2, nb, 2, nftau))
    for b in range(nb):
        for s in range(2):
            ftau[b, s, b, s, :] = f
    muimp = np.zeros((nb, 2, nb, 2))
    for b in range(nb):
        for s in range(2):
            muimp[b, s, b, s] = -mu       # engine convention: muimp = -mu (cthyb_solver.py:255)
    nf = 2 * nb
    q = dict(qmc)
    q["Nftau"] = nftau
    return Problem(name, nb, beta, ham, qns, np.ascontiguousarray(um8.reshape(nf, nf, nf, nf)),
                   muimp, ftau, q, eps, v, mu)


def make_problem(cfg: str, nftau: int = 5000, beta: float = None, **qmc_over) -> Problem:
    """The five named configurations of BASELINE.json (table in SURVEY.md §8(d)).  `beta` overrides the inverse temperature (tests
    of moves that are only accepted at high temperature); TaudiffMax is then capped at beta."""
    cfg = cfg.upper()
    if cfg == "C1":
        um = u_density(1, 2.0, 0.0, 0.0)
        p = _assemble("C1", 1, beta or 10.0, "Density", "Nt Szt Azt", um, 1.0, 8,
                      {"Nwarmups": 100000, "Nmeas": 10000, "NCorr": 100, "TaudiffMax": 10.0}, nftau)
    elif cfg == "C2":
        u, jj, v = 4.0, 0.75, 2.5
        um = u_density(2, u, v, jj)
        mu = 0.5 * (u + (2 - 1) * (2 * v - jj))
        p = _assemble("C2", 2, beta or 50.0, "Density", "Nt Szt Azt", um, mu, 8,
                      {"TaudiffMax": 5.0}, nftau)
    elif cfg == "C3":
        u, jj, v = 4.0, 0.6, 2.8
        um = u_kanamori(3, u, v, jj)
        mu = 0.5 * (u + (3 - 1) * (2 * v - jj))
        p = _assemble("C3", 3, beta or 40.0, "Kanamori", "Nt Szt Qzt", um, mu, 6,
                      {"TaudiffMax": 4.0}, nftau)
    elif cfg == "C4":
        um, cub = u_coulomb(5, 4.0, 8.615, 5.385)
        ud = udens_from_ufull(cub)
        mu = 0.5 * ud[0, 0].sum()
        p = _assemble("C4", 5, beta or 20.0, "Coulomb", "Nt Szt All", um, mu, 6,
                      {"TaudiffMax": 2.0, "Nmeas": 2000}, nftau)
    elif cfg == "C5":
        um = u_density(1, 2.0, 0.0, 0.0)
        p = _assemble("C5", 1, beta or 200.0, "Density", "Nt Szt Azt", um, 1.0, 8,
                      {"TaudiffMax": 3.0}, nftau)
    else:
        raise ValueError("unknown config %r" % cfg)
    if beta is not None:
        p.qmc["TaudiffMax"] = min(p.qmc["TaudiffMax"], float(beta))
    p.qmc.update(qmc_over)
    return p
