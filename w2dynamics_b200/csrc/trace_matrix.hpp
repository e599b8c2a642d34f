// Blocked local trace in the H_loc eigenbasis (quantum-number blocks larger than 1x1: configs C3, C4).
//
// Reference semantics (files under /root/reference/src/ctqmc_fortran):
//   get_Trace_EB / get_Trace_EB_fullmatrix   Trace.F90:5354-5549, 5734-6033  (ket/bra chains, split_log renormalisation)
//   update_trace_EB                          Trace.F90:6109-6209              (cache hand-over on accept)
//   update_b2_states_eb                      Trace.F90:6213-6469              (states at beta/2 for the density matrix)
//
// B200-first design (not the reference's per-operator `state/cache` pointers tied to the sliding window):
//   For the outer superstate of dimension m the trace is sum_j <j| P O_n P ... O_1 P |j>.  Every operator owns one cache
//   slot holding KET_i = O_i P ... O_1 P I_m (d_i x m, columns normalised, one log scale per column) and
//   BRA_i = (I_m P O_n ... P O_i)^T (d_{i-1} x m).  Two frontiers say which are valid: kets for indices < kv, bras for
//   indices >= bv.  A proposal touching positions [i1, i2] extends the frontiers lazily to i1-1 / i2 (work that is kept
//   whether or not the move is accepted), multiplies through the changed sub-string only and closes with one dot product
//   per outer state.  An accepted move just moves the frontiers (kv = i1, bv = i2 +- 2); nothing is copied.  This is exact
//   (same products as a from-scratch evaluation) and independent of the sliding window.
//   Slots are addressed through a per-operator slot id so that inserting / removing operators shifts 2-byte ids, not
//   matrices.  Slot storage is compact (leading dimension = actual block dimension).
//
//   The product itself, Y = O * diag(exp(-(E-E0) dt)) * X, runs on the FP64 tensor pipe: one warp issues
//   mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) over 8x8 output tiles.  X lives in the warp's shared-memory work matrix
//   (column-major, leading dimension ld == 4 mod 16 doubles, which makes the B-fragment loads bank-conflict free) and is
//   overwritten in place, column tile by column tile (column j of Y depends on column j of X only, and all row tiles of
//   a column tile are accumulated in registers before the first store).  The operator blocks are read through the
//   read-only path (a few MB for the 5-orbital problem, L1/L2 resident, shared by all chains).  Measured on the C4
//   workload 77 % of the trace flops sit in 28-wide outer blocks (61 kflop per move on average), so the dense path is
//   the common case, not the exception.
#pragma once
#include "warp.hpp"

namespace w2b {

struct Tables;
struct Chains;
struct Ctx;

constexpr double NEG_HUGE = -1.7976931348623157e308;
constexpr int MAT_DMAX = 32;          // largest block dimension of the in-place tensor path (4 row tiles in registers)

// leading dimension of the work matrix for blocks up to dmax: the smallest value >= dmax that is 4 mod 16
#if W2B_DEVICE
__host__ __device__
#endif
inline int mat_ld(int dmax) { int ld = 4; while (ld < dmax) ld += 16; return ld; }

// X = I_m, log scales 0
W2B_HDN void mat_identity(double* X, int ld, double* lg, int m) {
    for (int j = 0; j < m; ++j) for (int r = lane_id(); r < m; r += W2B_WARP) X[r + j * ld] = (r == j) ? 1.0 : 0.0;
    for (int j = lane_id(); j < m; j += W2B_WARP) lg[j] = 0.0;
    wsync();
}
W2B_HDN void mat_copy(double* dst, const double* src, int cnt) {
    for (int e = lane_id(); e < cnt; e += W2B_WARP) dst[e] = src[e];
}
// (d x m) column-major copies with different leading dimensions.  Lanes are laid out as P rows x 32/P columns with
// P the power of two >= d, so small blocks move several columns per instruction.
W2B_HD int pow2_rows(int d) { int p = 1; while (p < d && p < W2B_WARP) p <<= 1; return p; }
W2B_HDN void mat_move(double* dst, int ldd, const double* src, int lds, int d, int m) {
#if W2B_DEVICE
    const int P = pow2_rows(d), r = lane_id() & (P - 1), cj = lane_id() / P, cstep = W2B_WARP / P;
    if (r < d) for (int j = cj; j < m; j += cstep) dst[r + (size_t)j * ldd] = src[r + (size_t)j * lds];
#else
    for (int j = 0; j < m; ++j) for (int r = 0; r < d; ++r) dst[r + (size_t)j * ldd] = src[r + (size_t)j * lds];
#endif
}

// power-of-two renormalisation: returns the scale 2^-e, e = exponent of mx; lgadd = e ln 2
// (the reference divides by the largest element and adds its log, Trace.F90:5802-5822 / split_log 5341-5348; a power
// of two is exact and needs no log())
constexpr double LN2 = 0.6931471805599453094;
W2B_HD double renorm_pow2(double mx, double& lgadd) {
    int e;
#if W2B_DEVICE
    e = ilogb(mx);
    lgadd = (double)e * LN2;
    return scalbn(1.0, -e);
#else
    e = std::ilogb(mx);
    lgadd = (double)e * LN2;
    return std::scalbn(1.0, -e);
#endif
}

#if W2B_DEVICE
// D(8x8) += A(8x4, row) * B(4x8, col).  Fragments: lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2*(l%4) + {0,1}].
W2B_HD void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#endif

// X(d_in x m) <- O(d_out x d_in) * diag(exp(-(E_c - E_0) dt)) * X, in place; every column is renormalised by a power
// of two near its largest element and lg[j] += -(E_0 - Egs) dt + e ln2.  `pw` is warp scratch for the d_in propagation
// factors.  Returns the target block.
template <class TablesT>
W2B_HDN int mat_apply(const TablesT& T, int sst_in, int opc, double dt, double* X, int ld, double* lg, double* pw, int m) {
    const int d_in = ldg(&T.dim[sst_in]);
    const double* ev = T.evals + ldg(&T.offset[sst_in]);
    const double e0 = ldg(&ev[0]);
    const int sst_out = (int)ldg(&T.connect[sst_in * T.nops + opc]);
    const int d_out = ldg(&T.dim[sst_out]);
    const double* O = T.opdata + ldg(&T.op_off[(size_t)sst_in * T.nops + opc]);
    const double shift = -(e0 - T.egs) * dt;
    if (d_in == 1 && d_out == 1) {                               // scalar link of the chain: one multiply per outer state
        const double o = ldg(&O[0]);
        for (int j = lane_id(); j < m; j += W2B_WARP) {
            const double y = o * X[(size_t)j * ld];
            if (y != 0.0) { double la; const double sc = renorm_pow2(fabs(y), la); X[(size_t)j * ld] = y * sc; lg[j] += shift + la; }
            else X[(size_t)j * ld] = 0.0;
        }
        wsync();
        return sst_out;
    }
    for (int c = lane_id(); c < d_in; c += W2B_WARP) pw[c] = c ? w_exp(-(ldg(&ev[c]) - e0) * dt) : 1.0;
    wsync();
#if W2B_DEVICE
    const int lane = lane_id(), g = lane >> 2, t = lane & 3;
    const int ntiles = (m + 7) >> 3, mtiles = (d_out + 7) >> 3, ksteps = (d_in + 3) >> 2;
    for (int nt = 0; nt < ntiles; ++nt) {
        double acc[4][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) { acc[mt][0] = 0.0; acc[mt][1] = 0.0; }
        const int n = nt * 8 + g;
        const bool nok = n < m;
        const double* xcol = X + (size_t)n * ld;
        for (int ks = 0; ks < ksteps; ++ks) {
            const int k = ks * 4 + t;
            const bool kok = k < d_in;
            const double b = (kok && nok) ? xcol[k] * pw[k] : 0.0;
            const double* ocol = O + (size_t)k * d_out;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                if (mt < mtiles) {                                 // warp-uniform
                    const int r = mt * 8 + g;
                    const double a = (kok && r < d_out) ? ldg(&ocol[r]) : 0.0;
                    dmma884(acc[mt][0], acc[mt][1], a, b);
                }
            }
        }
        wsync();                                                   // every lane has consumed its part of this column tile
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int col = nt * 8 + 2 * t + q;
            double mx = 0.0;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) mx = fmax(mx, fabs(acc[mt][q]));
            // column maximum over the 8 lanes sharing t: order-preserving compare of the high words is enough for the exponent
            int hi = __double2hiint(mx);
            hi = max(hi, __shfl_xor_sync(W2B_FULL, hi, 4)); hi = max(hi, __shfl_xor_sync(W2B_FULL, hi, 8)); hi = max(hi, __shfl_xor_sync(W2B_FULL, hi, 16));
            const int ef = hi >> 20;
            double sc = 1.0, la = 0.0; bool nonzero = true;
            if (wany(ef == 0 || ef > 2040)) {                      // zero / subnormal / huge column somewhere: exact slow path
                double mxx = mx;
                mxx = fmax(mxx, __shfl_xor_sync(W2B_FULL, mxx, 4)); mxx = fmax(mxx, __shfl_xor_sync(W2B_FULL, mxx, 8)); mxx = fmax(mxx, __shfl_xor_sync(W2B_FULL, mxx, 16));
                if (mxx > 0.0) sc = renorm_pow2(mxx, la); else nonzero = false;
            } else { const int e = ef - 1023; sc = __hiloint2double((1023 - e) << 20, 0); la = (double)e * LN2; }
            if (col < m) {
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) { const int r = mt * 8 + g; if (mt < mtiles && r < d_out) X[r + (size_t)col * ld] = acc[mt][q] * sc; }
                if (g == 0 && nonzero) lg[col] += shift + la;
            }
        }
        wsync();
    }
#else
    for (int j = 0; j < m; ++j) {
        double* xc = X + (size_t)j * ld;
        double y[MAT_DMAX]; double mx = 0.0;
        for (int r = 0; r < d_out; ++r) {
            double acc = 0.0;
            for (int c = 0; c < d_in; ++c) acc += O[r + (size_t)c * d_out] * (pw[c] * xc[c]);
            y[r] = acc; mx = std::fmax(mx, std::fabs(acc));
        }
        double sc = 1.0, la = 0.0;
        if (mx > 0.0) { sc = renorm_pow2(mx, la); lg[j] += shift + la; }
        for (int r = 0; r < d_out; ++r) xc[r] = y[r] * sc;
    }
#endif
    return sst_out;
}

// Close ket K (block sst, d x m, leading dimension ld, scales klg) with bra B (same block, compact d x m; nullptr = unit
// vectors e_j at tau = beta) across a propagation of length dt:  tr_j = sum_c B[c,j] exp(-(E_c-E_0) dt) K[c,j].
// pt receives (log, sign) per outer state, (tlog, tsign) the signed log-sum (addLogTr, Trace.F90:298-309).
template <class TablesT>
W2B_HDN void mat_close(const TablesT& T, int sst, double dt, const double* K, int ld, const double* klg, const double* B, const double* blg,
                      double* pw, int m, double* pt, double& tlog, double& tsign) {
    const int d = ldg(&T.dim[sst]);
    const double* ev = T.evals + ldg(&T.offset[sst]);
    const double e0 = ldg(&ev[0]);
    for (int c = lane_id(); c < d; c += W2B_WARP) pw[c] = c ? w_exp(-(ldg(&ev[c]) - e0) * dt) : 1.0;
    wsync();
    for (int j = lane_id(); j < m; j += W2B_WARP) {
        double s;
        if (B) { s = 0.0;
#pragma unroll 4
            for (int c = 0; c < d; ++c) s += B[c + (size_t)j * d] * pw[c] * K[c + (size_t)j * ld]; }
        else s = pw[j] * K[j + (size_t)j * ld];
        if (s != 0.0) { pt[2 * j] = klg[j] + (B ? blg[j] : 0.0) - (e0 - T.egs) * dt + w_log(fabs(s)); pt[2 * j + 1] = s < 0.0 ? -1.0 : 1.0; }
        else { pt[2 * j] = NEG_HUGE; pt[2 * j + 1] = 0.0; }
    }
    wsync();
    if (m == 1) { tlog = pt[0]; tsign = pt[1]; return; }
    // signed log-sum over the outer states, lane-parallel
    double mx = NEG_HUGE;
    for (int j = lane_id(); j < m; j += W2B_WARP) if (pt[2 * j + 1] != 0.0) mx = fmax(mx, pt[2 * j]);
    mx = wmaxd(mx);
    if (mx == NEG_HUGE) { tlog = NEG_HUGE; tsign = 0.0; return; }
    double sum = 0.0;
    for (int j = lane_id(); j < m; j += W2B_WARP) if (pt[2 * j + 1] != 0.0) sum += pt[2 * j + 1] * w_exp(pt[2 * j] - mx);
    sum = wsumd(sum);
    if (sum == 0.0) { tlog = NEG_HUGE; tsign = 0.0; return; }
    tlog = mx + w_log(fabs(sum)); tsign = sum < 0.0 ? -1.0 : 1.0;
}

}  // namespace w2b
