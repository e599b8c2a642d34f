"""ORACLE (test infrastructure, not product code) -- atomic-problem setup.

CPU restatement (numpy + scipy LAPACK) of the reference's ``init_solver`` chain
(``src/ctqmc_fortran/CTQMC.F90:7894-8068``):

* Fock space / quantum-number blocks: ``States.F90:58-184`` (+ QN functions 188-325)
* occupation-basis c/c^dagger blocks and ``Connect``: ``Operator.F90:115-197``
* H_loc: ``oneparticle_hamiltonian`` Operator.F90:482-542, ``u_hamiltonian``
  1042-1108, ``force_hermitian`` 1110-1124
* automatic block refinement: ``analyze_hamiltonian`` Operator.F90:303-442
* DSYEV per block Operator.F90:637-670, ``set_HEVec`` 771-801 (EPSEVEC zeroing and
  the *non-sqrt* renormalisation v/sum(v^2)), ``transform_Psis`` 1130-1168

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may
import this module.  Parity status: the reference's own tests pin nothing on this
path -- "parity unpinned" by the reference; pins are the analytic
limits and block-structure table of SURVEY.md §8 checked in tests/test_oracle_atomic.py.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from scipy.linalg import lapack


def _bit(b, s, nb):
    """bit index of flavour (band b, spin s), 0-based: (b) + nb*(s)  (Operator.F90:160)."""
    return b + nb * s


def _qn_values(nb, qns):
    """Quantum-number tuple for every Fock state (States.F90:188-325)."""
    n = 4 ** nb
    st = np.arange(n)
    occ = ((st[:, None] >> np.arange(2 * nb)[None, :]) & 1).astype(np.int64)  # bit i
    cols = []
    if "Nt" in qns:
        cols.append(occ.sum(1).astype(float))
    if "Szt" in qns:
        cols.append(0.5 * (occ[:, nb:].sum(1) - occ[:, :nb].sum(1)))
    if "Qzt" in qns:
        single = occ[:, :nb] ^ occ[:, nb:]
        cols.append((single * (10 ** np.arange(nb))[None, :]).sum(1).astype(float))
    if "Azt" in qns:
        cols.append((occ * (10.0 ** np.arange(2 * nb))[None, :]).sum(1))
    if "Lzt" in qns or "Jzt" in qns:
        lz = np.tile(np.arange(1, nb + 1) - (nb + 1) / 2.0, 2)
        lzt = (occ * lz[None, :]).sum(1)
        if "Lzt" in qns:
            cols.append(lzt)
        if "Jzt" in qns:
            cols.append(lzt + 0.5 * (occ[:, nb:].sum(1) - occ[:, :nb].sum(1)))
    if not cols:
        return np.zeros((n, 1))
    return np.stack(cols, 1)


def qns2substates(nb, qns):
    """States.F90:149-184: blocks numbered by their lowest member state."""
    q = _qn_values(nb, qns)
    n = q.shape[0]
    s2s = -np.ones(n, dtype=np.int64)
    nsub = 0
    for st1 in range(n):
        if s2s[st1] != -1:
            continue
        same = np.all(np.abs(q - q[st1][None, :]) < 1e-12, axis=1) & (s2s == -1)
        same[:st1] = False
        s2s[same] = nsub
        nsub += 1
    return nsub, s2s


@dataclass
class Atomic:
    nbands: int
    nstates: int
    nsst: int
    states2sst: np.ndarray          # (nstates,) block of each Fock state
    block_states: list              # per block: ascending Fock-state numbers
    offset: np.ndarray              # (nsst,)
    dim: np.ndarray                 # (nsst,)
    connect: np.ndarray             # (nsst, nb, 2, 2) -> target block or -1 ; last index ca-1 (0=c^dag,1=c)
    eval: list                      # per block eigenvalues ascending
    evec: list                      # per block (d,d) columns = eigenvectors (after set_HEVec)
    psis_eb: dict                   # (sst,b,s,ca) -> (d_target, d_sst) ndarray
    psis_occ: dict
    hloc: list
    egs: float

    @property
    def nstatesmax(self):
        return int(self.dim.max())


def _fock_ops(nb):
    """Dense full-Fock-space creation operators with the reference's sign
    convention: (-1)^(occupied bits below), Operator.F90:158-172."""
    n = 4 ** nb
    ops = []
    st = np.arange(n)
    for s in range(2):
        for b in range(nb):
            i = _bit(b, s, nb)
            mask = (1 << i) - 1
            par = np.array([bin(x & mask).count("1") for x in st])
            sig = 1.0 - 2.0 * (par % 2)
            src = st[(st >> i) & 1 == 0]
            cd = np.zeros((n, n))
            cd[src | (1 << i), src] = sig[src]
            ops.append(cd)
    # index by bit position
    return ops


def _hloc_full(nb, u_matrix, muimp):
    """Full-Fock H = sum muimp c^+ c + 1/2 sum U_1234 c1^+ c2^+ c4 c3
    (Operator.F90:482-542, 1042-1108).  Taking diagonal blocks afterwards equals the
    reference's block-wise construction."""
    from scipy import sparse
    cd = [sparse.csr_matrix(m) for m in _fock_ops(nb)]
    c = [m.T.tocsr() for m in cd]
    n = 4 ** nb
    h = sparse.csr_matrix((n, n))

    def fl(b, s):
        return _bit(b, s, nb)
    for b1 in range(nb):
        for s1 in range(2):
            for b2 in range(nb):
                for s2 in range(2):
                    val = muimp[b1, s1, b2, s2]
                    if val != 0:
                        h = h + val * (cd[fl(b1, s1)] @ c[fl(b2, s2)])
    nf = 2 * nb
    nz = np.argwhere(u_matrix != 0)
    for i1, i2, i3, i4 in nz:
        b1, s1 = divmod(i1, 2)
        b2, s2 = divmod(i2, 2)
        b3, s3 = divmod(i3, 2)
        b4, s4 = divmod(i4, 2)
        uval = 0.5 * u_matrix[i1, i2, i3, i4]
        h = h + uval * (cd[fl(b1, s1)] @ cd[fl(b2, s2)] @ c[fl(b4, s4)] @ c[fl(b3, s3)])
    return np.asarray(h.todense())


def _blocks_from_map(s2s, nsub):
    return [np.nonzero(s2s == k)[0] for k in range(nsub)]


def analyze_hamiltonian(nb, hfull, s2s_in, nsub_in, threshold):
    """Operator.F90:303-442: split every QN block into H-connected components (numbered in
    order of lowest contained block-local index), then merge target blocks until every
    c/c^dagger maps a block into a single block."""
    n = hfull.shape[0]
    s2s = -np.ones(n, dtype=np.int64)
    nsub = -1
    for blk in _blocks_from_map(s2s_in, nsub_in):
        hb = hfull[np.ix_(blk, blk)]
        adj = (hb != 0.0) & (np.abs(hb) > threshold)
        d = len(blk)
        local = -np.ones(d, dtype=np.int64)
        while True:
            free = np.nonzero(local == -1)[0]
            if free.size == 0:
                break
            nsub += 1
            members = [int(free[0])]
            local[free[0]] = nsub
            k = 0
            while k < len(members):
                node = members[k]
                for edge in np.nonzero(adj[node])[0]:
                    if local[edge] == -1:
                        local[edge] = nsub
                        members.append(int(edge))
                k += 1
        s2s[blk] = local
    nsub += 1
    # merge step (literal restatement incl. renumbering)
    st = np.arange(n)
    change = True
    while change:
        change = False
        for b in range(nb):
            for s in range(2):
                for ca in (1, 2):
                    i = _bit(b, s, nb)
                    occ = (st >> i) & 1
                    if ca == 1:
                        valid = occ == 0
                        img_state = st | (1 << i)
                    else:
                        valid = occ == 1
                        img_state = st & ~(1 << i)
                    cimg = np.where(valid, s2s[img_state], -1)
                    # first (idx1<idx2) in same block with different valid images
                    v = np.nonzero(valid)[0]
                    blk = s2s[v]
                    img = cimg[v]
                    mn = np.full(nsub, n + 1, dtype=np.int64)
                    mx = np.full(nsub, -1, dtype=np.int64)
                    np.minimum.at(mn, blk, img)
                    np.maximum.at(mx, blk, img)
                    conflict = (mx >= 0) & (mn != mx)
                    if not conflict.any():
                        continue
                    idx1 = v[conflict[blk]][0]
                    sel = (v > idx1) & (blk == s2s[idx1]) & (img != cimg[idx1])
                    found = (idx1, v[sel][0])
                    big = max(cimg[found[0]], cimg[found[1]])
                    small = min(cimg[found[0]], cimg[found[1]])
                    was_big = s2s == big
                    above = s2s > big
                    s2s[was_big] = small
                    s2s[above] -= 1
                    change = True
                    break
                if change:
                    break
            if change:
                break
    nsub = int(s2s.max()) + 1
    return nsub, s2s


def build_atomic(nbands, qns, u_matrix, muimp_full, epsblock=1e-13, epsevec=1e-15) -> Atomic:
    """init_solver up to set_EB_Psis (CTQMC.F90:7957-8031)."""
    nb = nbands
    n = 4 ** nb
    u_matrix = np.asarray(u_matrix, dtype=float).reshape((2 * nb,) * 4)
    muimp_full = np.asarray(muimp_full, dtype=float).reshape(nb, 2, nb, 2)
    nsub, s2s = qns2substates(nb, qns)
    hfull = _hloc_full(nb, u_matrix, muimp_full)
    if "All" in qns:
        thr = epsblock * max(np.abs(muimp_full).max(), np.abs(u_matrix).max())
        # the reference analyses the *block-diagonal, hermitised* H of the first pass
        hsym = 0.5 * (hfull + hfull.T)
        mask = s2s[:, None] == s2s[None, :]
        nsub, s2s = analyze_hamiltonian(nb, np.where(mask, hsym, 0.0), s2s, nsub, thr)
    blocks = _blocks_from_map(s2s, nsub)
    dim = np.array([len(b) for b in blocks])
    offset = np.concatenate([[0], np.cumsum(dim)[:-1]])

    cd_full = _fock_ops(nb)
    connect = -np.ones((nsub, nb, 2, 2), dtype=np.int64)
    psis_occ = {}
    for b in range(nb):
        for s in range(2):
            cd = cd_full[_bit(b, s, nb)]
            for ca, op in ((1, cd), (2, cd.T)):
                for k, blk in enumerate(blocks):
                    cols = op[:, blk]
                    tgt_states = np.nonzero(np.any(cols != 0, axis=1))[0]
                    if tgt_states.size == 0:
                        continue
                    tgts = np.unique(s2s[tgt_states])
                    # Operator.F90:145-197: the *last* connected block in loop order wins;
                    # after analyze_hamiltonian the target is unique
                    t = int(tgts.max())
                    connect[k, b, s, ca - 1] = t
                    psis_occ[(k, b, s, ca)] = op[np.ix_(blocks[t], blk)].copy()

    hloc, evals, evecs = [], [], []
    for blk in blocks:
        hb = hfull[np.ix_(blk, blk)].copy()
        # force_hermitian (Operator.F90:1110-1124)
        hb = 0.5 * (hb + hb.T)
        hloc.append(hb)
        w, v, info = lapack.dsyev(hb, compute_v=1, lower=0)
        assert info == 0
        v = v.copy()
        v[np.abs(v) < epsevec] = 0.0
        v = v / (v ** 2).sum(0)[None, :]          # quirk: v/sum(v^2), Operator.F90:788-789
        evals.append(w)
        evecs.append(v)
    psis_eb = {}
    for (k, b, s, ca), m in psis_occ.items():
        t = connect[k, b, s, ca - 1]
        psis_eb[(k, b, s, ca)] = evecs[t].T @ (m @ evecs[k])
    egs = min(w[0] for w in evals)
    return Atomic(nb, n, nsub, s2s, blocks, offset, dim, connect, evals, evecs, psis_eb,
                  psis_occ, hloc, float(egs))


def pack_tables(at: Atomic):
    """Flatten the static tables for the C++ chain oracle (oracle/chain.cpp)."""
    nsst, nb = at.nsst, at.nbands
    nops = nb * 2 * 2
    op_off = -np.ones((nsst, nops), dtype=np.int64)
    chunks = []
    pos = 0
    for k in range(nsst):
        for b in range(nb):
            for s in range(2):
                for ca in (1, 2):
                    key = (k, b, s, ca)
                    if key in at.psis_eb:
                        m = np.asfortranarray(at.psis_eb[key])   # column-major (d_t, d_k)
                        op_off[k, (b * 2 + s) * 2 + (ca - 1)] = pos
                        chunks.append(m.ravel(order="F"))
                        pos += m.size
    opdata = np.concatenate(chunks) if chunks else np.zeros(0)
    eval_flat = np.concatenate(at.eval)
    return dict(nsst=nsst, nb=nb, dim=at.dim.astype(np.int32), offset=at.offset.astype(np.int32),
                connect=np.ascontiguousarray(at.connect.reshape(nsst, nops)).astype(np.int32),
                op_off=op_off, opdata=np.ascontiguousarray(opdata),
                evals=np.ascontiguousarray(eval_flat), egs=at.egs)
