"""Exact reference for small Anderson impurity models (test infrastructure; numpy / scipy only, no oracle, no engine code):
the impurity of one of the named configurations coupled to a bath of n_bath levels per flavour, diagonalised densely in the full
Fock space of impurity + bath.  Gives G(tau) = <c_f(tau) c_f^+(0)> and <n_i n_j> at inverse temperature beta -- the answer a CT-HYB
run of the same model has to reproduce inside its error bars, for ANY interaction (Kanamori spin-flip / pair-hopping included)."""
import numpy as np
import scipy.sparse as sp

from w2dynamics_b200 import problems


def small_bath_problem(cfg, n_bath, beta, nftau=2001, **qmc):
    """The named configuration's impurity (same U-matrix, mu, quantum numbers) with a Bethe bath of n_bath levels per flavour."""
    p = problems.make_problem(cfg, nftau=nftau, beta=beta, **qmc)
    eps, v = problems.bethe_bath(n_bath)
    f = problems.hybr_tau(eps, v, beta, nftau)
    p.ftau[...] = 0.0
    for b in range(p.nbands):
        for s in range(2):
            p.ftau[b, s, b, s, :] = f
    p.eps_bath, p.v_bath = eps, v
    return p


def kanamori2_problem(n_bath, beta, jj=0.6, nftau=2001, **qmc):
    """Two orbitals with the full Kanamori interaction (spin flip and pair hopping, 2 x 2 blocks -> the blocked trace path) at half
    filling.  NOTE on the bath: with ONE level per flavour the bath can hold one electron per flavour, so every configuration whose
    flavour-f operators do not alternate has zero weight although its local trace is finite for J != 0 -- the pair / 4-operator
    moves are then NOT ergodic (measured: G(tau -> 0) 1.6 % low at J = 0.6, 2.5 % at J = 1.2, with <k> and all occupations exact);
    two levels per flavour remove that and the same chains agree with the exact result.  Use n_bath >= 2 for J != 0."""
    u, v = 4.0, 2.8
    mu = 0.5 * (u + (2 * v - jj))
    q = {"TaudiffMax": min(4.0, beta)}
    q.update(qmc)
    p = problems._assemble("K2", 2, beta, "Kanamori", "Nt Szt Qzt", problems.u_kanamori(2, u, v, jj), mu, n_bath, q, nftau)
    eps, vb = problems.bethe_bath(n_bath)
    f = problems.hybr_tau(eps, vb, beta, nftau)
    p.ftau[...] = 0.0
    for b in range(2):
        for s in range(2):
            p.ftau[b, s, b, s, :] = f
    p.eps_bath, p.v_bath = eps, vb
    return p


def offdiag_two_orbital_problem(beta, nftau=2001, kanamori=False, **qmc):
    """Two orbitals, density-density (or, `kanamori`, full Kanamori) interaction, and a bath whose levels couple to BOTH orbitals: three levels per spin, level q hops to
    orbital 0 with V[0, q] and to orbital 1 with V[1, q] -> orbital-off-diagonal hybridisation F_{(0 s), (1 s)}(tau) != 0 (`offdiag = 1`).
    F_ab(tau) = sum_q V_aq V_bq / (exp(-tau eps_q) + exp((beta - tau) eps_q)), the off-diagonal form of problems.hybr_tau.
    NOTE on ergodicity: with spin-flip / pair-hopping terms H_loc does not conserve the particle number per flavour, configurations like
    c+_{0 up} c+_{0 dn} c_{1 dn} c_{1 up} carry weight through F_01, and single pair insertions cannot reach them (the quantum-number check
    rejects one flavour-changing pair): pair moves alone are then NOT ergodic (measured: G_01 comes out 3-5 x too large while the diagonal
    G and all occupations agree) -- the reference's 4-operator / flavour-change moves exist for this.  Density-density interactions
    conserve every flavour's particle number and pair moves are ergodic."""
    u, v, jj = 4.0, 2.8, 0.6
    mu = 0.5 * (u + (2 * v - jj))
    q = {"TaudiffMax": min(4.0, beta), "offdiag": 1, "PercentageGlobalMove": 0.0}
    q.update(qmc)
    if kanamori:
        p = problems._assemble("K2o", 2, beta, "Kanamori", "Nt Szt Qzt", problems.u_kanamori(2, u, v, jj), mu, 3, q, nftau)
    else:
        p = problems._assemble("D2o", 2, beta, "Density", "Nt Szt Azt", problems.u_density(2, u, v, jj), mu, 3, q, nftau)
    eps = np.array([-0.8, 0.15, 1.1])
    vq = np.array([[0.65, 0.5, 0.45], [0.4, -0.55, 0.6]])                 # [orbital, level]
    tau = np.linspace(0.0, beta, nftau)
    den = np.exp(-tau[:, None] * eps[None, :]) + np.exp((beta - tau[:, None]) * eps[None, :])
    p.ftau[...] = 0.0
    for s in range(2):
        for b1 in range(2):
            for b2 in range(2):
                p.ftau[b1, s, b2, s, :] = ((vq[b1] * vq[b2])[None, :] / den).sum(-1)
    # general bath for the exact solver: modes (spin, level); hopping matrix [impurity flavour 2 b + s, bath mode]
    nmode = 2 * eps.size
    p.bath_eps = np.concatenate([eps, eps])
    p.bath_v = np.zeros((4, nmode))
    for s in range(2):
        for b in range(2):
            p.bath_v[2 * b + s, s * eps.size:(s + 1) * eps.size] = vq[b]
    return p


def solve_general(p, pairs, nbins=50):
    """Exact G_ab(tau) = <c_a(tau) c_b^+(0)> (bin averages) and <n_i n_j> for a GENERAL bath (p.bath_eps[m], p.bath_v[flavour, m])."""
    nb, nf, beta = p.nbands, 2 * p.nbands, p.beta
    eps, v = np.asarray(p.bath_eps), np.asarray(p.bath_v)
    nm = nf + eps.size
    c = _annihilators(nm)
    cd = [m.T.tocsr() for m in c]
    u = np.asarray(p.u_matrix).reshape((nf,) * 4)
    mu = np.asarray(p.muimp).reshape(nf, nf)
    h = sp.csr_matrix((1 << nm, 1 << nm))
    for i in range(nf):
        for j in range(nf):
            if mu[i, j] != 0.0:
                h = h + mu[i, j] * (cd[i] @ c[j])
    for i, j, k, l in np.argwhere(np.abs(u) > 1e-14):
        h = h + 0.5 * u[i, j, k, l] * (cd[i] @ cd[j] @ (c[l] @ c[k]))
    for m in range(eps.size):
        h = h + eps[m] * (cd[nf + m] @ c[nf + m])
        for i in range(nf):
            if v[i, m] != 0.0:
                h = h + v[i, m] * (cd[i] @ c[nf + m] + cd[nf + m] @ c[i])
    hd = h.toarray()
    assert np.abs(hd - hd.T).max() < 1e-12
    w, vec = np.linalg.eigh(hd)
    w -= w.min()
    bw = np.exp(-beta * w)
    z = bw.sum()
    edges = np.linspace(0.0, beta, nbins + 1)
    cmat = {f: vec.T @ (cd[f] @ vec) for f in {x for pr in pairs for x in pr}}       # <m| c_f^+ |n>
    g = np.zeros((len(pairs), nbins))
    for a, (fa, fb) in enumerate(pairs):
        m2 = cmat[fa] * cmat[fb]                                        # <n|c_a|m> <m|c_b^+|n> (real matrices)
        im, in_ = np.nonzero((np.abs(m2) > 1e-15) & (np.maximum(bw[:, None], bw[None, :]) / z > 1e-13))
        mm, em, en = m2[im, in_], w[im], w[in_]
        de = em - en
        small = np.abs(de) < 1e-9
        for ib in range(nbins):
            t1, t2 = edges[ib], edges[ib + 1]
            e1, e2 = np.exp(-(beta - t1) * en - t1 * em), np.exp(-(beta - t2) * en - t2 * em)
            with np.errstate(divide="ignore", invalid="ignore"):
                avg = np.where(small, e1, (e1 - e2) / (de * (t2 - t1)))
            g[a, ib] = (mm * avg).sum() / z
    keep = bw / z > 1e-14
    rho_diag = ((vec[:, keep] ** 2) @ bw[keep]) / z
    st = np.arange(1 << nm)
    occ = np.array([[(rho_diag * ((st >> i) & 1) * ((st >> j) & 1)).sum() for j in range(nf)] for i in range(nf)])
    return g, occ.reshape(nb, 2, nb, 2)


_CACHE = {}


def solve_cached(key, p, **kw):
    if key not in _CACHE:
        _CACHE[key] = solve(p, **kw)
    return _CACHE[key]


def _annihilators(nmodes):
    dim = 1 << nmodes
    st = np.arange(dim)
    ops = []
    for i in range(nmodes):
        occ = (st >> i) & 1
        src = st[occ == 1]
        par = np.zeros(src.size, dtype=np.int64)
        for j in range(i):
            par += (src >> j) & 1
        ops.append(sp.csr_matrix(((-1.0) ** par, (src & ~(1 << i), src)), shape=(dim, dim)))
    return ops


def solve(p, nbins=50, flavours=(0,), nleg=12, want_rho=False, niw=0):
    """Returns (gbin[len(flavours), nbins] = EXACT averages of G(tau) over nbins equal tau bins, occ[nb,2,nb,2],
    gleg[len(flavours), nleg] = sqrt(2l+1) int_0^beta P_l(2 tau / beta - 1) G(tau) dtau in closed form).  Mode index: impurity
    flavour i = 2 b + s -> i; bath level q of flavour i -> nf + i * n_bath + q."""
    from scipy.special import ive
    nb, nf, beta = p.nbands, 2 * p.nbands, p.beta
    eps, v = np.asarray(p.eps_bath), np.asarray(p.v_bath)
    nbath = eps.size
    nm = nf * (1 + nbath)
    c = _annihilators(nm)
    cd = [m.T.tocsr() for m in c]
    u = np.asarray(p.u_matrix).reshape((nf,) * 4)
    mu = np.asarray(p.muimp).reshape(nf, nf)
    h = sp.csr_matrix((1 << nm, 1 << nm))
    for i in range(nf):
        for j in range(nf):
            if mu[i, j] != 0.0:
                h = h + mu[i, j] * (cd[i] @ c[j])
    pair = {}
    for i, j, k, l in np.argwhere(np.abs(u) > 1e-14):
        if (k, l) not in pair:
            pair[(k, l)] = c[l] @ c[k]
        h = h + 0.5 * u[i, j, k, l] * (cd[i] @ cd[j] @ pair[(k, l)])
    for i in range(nf):
        for q in range(nbath):
            m = nf + i * nbath + q
            h = h + eps[q] * (cd[m] @ c[m]) + v[q] * (cd[i] @ c[m] + cd[m] @ c[i])
    hd = h.toarray()
    assert np.abs(hd - hd.T).max() < 1e-12
    w, vec = np.linalg.eigh(hd)
    w -= w.min()
    bw = np.exp(-beta * w)
    z = bw.sum()
    edges = np.linspace(0.0, beta, nbins + 1)
    g = np.zeros((len(flavours), nbins))
    gleg = np.zeros((len(flavours), nleg))
    giw = np.zeros((len(flavours), 2 * niw), dtype=complex)
    for a, f in enumerate(flavours):
        m2 = (vec.T @ (cd[f] @ vec)) ** 2                             # |<m| c_f^+ |n>|^2
        # G(tau) = 1/Z sum_nm |<m|c+|n>|^2 exp(-(beta - tau) E_n - tau E_m): at tau -> beta the FINAL state's thermal weight counts
        im, in_ = np.nonzero((m2 > 1e-15) & (np.maximum(bw[:, None], bw[None, :]) / z > 1e-13))
        mm, em, en = m2[im, in_], w[im], w[in_]
        de = em - en
        small = np.abs(de) < 1e-9
        for ib in range(nbins):                                       # exact bin average; every exponent is <= 0
            t1, t2 = edges[ib], edges[ib + 1]
            e1, e2 = np.exp(-(beta - t1) * en - t1 * em), np.exp(-(beta - t2) * en - t2 * em)
            with np.errstate(divide="ignore", invalid="ignore"):
                avg = np.where(small, e1, (e1 - e2) / (de * (t2 - t1)))
            g[a, ib] = (mm * avg).sum() / z
        # Legendre coefficients: int_0^beta P_l(x) exp(-(beta - tau) E_n - tau E_m) dtau = beta (-1)^l i_l(a) exp(-beta (E_n + E_m) / 2),
        # a = beta (E_m - E_n) / 2, i_l = modified spherical Bessel function of the first kind (exponentially scaled for stability)
        aa = 0.5 * beta * de
        pref = mm * np.exp(-0.5 * beta * (en + em) + np.abs(aa)) / z          # exp(-beta min(E_n, E_m)) <= 1
        for l in range(nleg):
            with np.errstate(divide="ignore", invalid="ignore"):
                il = np.where(np.abs(aa) < 1e-12, 1.0 if l == 0 else 0.0, np.sqrt(np.pi / (2.0 * np.abs(aa))) * ive(l + 0.5, np.abs(aa)) * np.sign(aa) ** l)
            gleg[a, l] = np.sqrt(2 * l + 1.0) * beta * (-1.0) ** l * (pref * il).sum()
        if niw:
            # int_0^beta exp(i w_n tau) G(tau) dtau = -1/Z sum |<m|c+|n>|^2 (exp(-beta E_m) + exp(-beta E_n)) / (i w_n - (E_m - E_n)):
            # what MeasGiw accumulates ("giw-meas"; the physical G(i w) is its negative, cthyb_solver.py:814)
            wn = (2.0 * np.arange(-niw, niw) + 1.0) * np.pi / beta
            wt = mm * (bw[im] + bw[in_]) / z
            giw[a] = -(wt[None, :] / (1j * wn[:, None] - de[None, :])).sum(1)
    keep = bw / z > 1e-14
    rho_diag = ((vec[:, keep] ** 2) @ bw[keep]) / z                   # thermal weight of every occupation-number state
    st = np.arange(1 << nm)
    occ = np.array([[(rho_diag * ((st >> i) & 1) * ((st >> j) & 1)).sum() for j in range(nf)] for i in range(nf)])
    if niw:
        return g, occ.reshape(nb, 2, nb, 2), gleg, giw
    if not want_rho:
        return g, occ.reshape(nb, 2, nb, 2), gleg
    # one- and two-particle reduced density matrices of the impurity: rho1[i, j] = <c+_i c_j>,
    # rho2[i, j, k, l] = <c+_i c+_j c_k c_l>  (Meas_rho1 / rho2_from_Densitymatrix, CTQMC.F90:2884-3017: rho2(f4, f3, f2, f1) = Tr[rho c+_4 c+_3 c_2 c_1])
    vk, wk = vec[:, keep], bw[keep] / z

    def expect(op):
        return float((((op @ vk) * vk).sum(0) * wk).sum())
    rho1 = np.array([[expect(cd[i] @ c[j]) for j in range(nf)] for i in range(nf)])
    rho2 = np.zeros((nf,) * 4)
    for k in range(nf):
        for l in range(nf):
            if k == l:
                continue
            ckl = c[k] @ c[l]
            for i in range(nf):
                for j in range(nf):
                    if i != j:
                        rho2[i, j, k, l] = expect(cd[i] @ cd[j] @ ckl)
    return g, occ.reshape(nb, 2, nb, 2), gleg, rho1, rho2
