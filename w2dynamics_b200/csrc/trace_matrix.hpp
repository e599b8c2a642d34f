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
//   matrices.  Storage is compact (leading dimension = actual block dimension).
#pragma once
#include "warp.hpp"

namespace w2b {

struct Tables;
struct Chains;
struct Ctx;

constexpr double NEG_HUGE = -1.7976931348623157e308;

// X = I_m, log scales 0
W2B_HDN void mat_identity(double* X, double* lg, int m) {
    for (int e = lane_id(); e < m * m; e += W2B_WARP) X[e] = (e / m == e % m) ? 1.0 : 0.0;
    for (int j = lane_id(); j < m; j += W2B_WARP) lg[j] = 0.0;
    wsync();
}
W2B_HDN void mat_copy(double* dst, const double* src, int cnt) {
    for (int e = lane_id(); e < cnt; e += W2B_WARP) dst[e] = src[e];
}

// Y(d_out x m) = O(d_out x d_in) * diag(exp(-(E_c - E_0) dt)) * X(d_in x m); every column is renormalised by a power
// of two near its largest element and lg[j] += -(E_0 - Egs) dt + e ln2   (the reference divides by the largest element and
// adds its log, Trace.F90:5802-5822 / split_log 5341-5348; a power of two is exact and needs no log()).
// `pw` is warp scratch for the d_in propagation factors.  X is overwritten by the propagated input.  Returns the target block.
constexpr double LN2 = 0.6931471805599453094;
W2B_HD double renorm_pow2(double mx, double& lgadd) {          // returns the scale 2^-e, e = exponent of mx
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

template <class TablesT>
W2B_HDN int mat_apply(const TablesT& T, int sst_in, int opc, double dt, double* X, double* Y, double* lg, double* pw, int m) {
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
            const double y = o * X[j];
            if (y != 0.0) { double la; const double sc = renorm_pow2(fabs(y), la); Y[j] = y * sc; lg[j] += shift + la; }
            else Y[j] = 0.0;
        }
        wsync();
        return sst_out;
    }
    if (d_in > 1) {
        for (int c = lane_id(); c < d_in; c += W2B_WARP) pw[c] = c ? w_exp(-(ldg(&ev[c]) - e0) * dt) : 1.0;
        wsync();
    }
    for (int e = lane_id(); e < d_out * m; e += W2B_WARP) {
        const int j = e / d_out, r = e - j * d_out;
        const double* xc = X + (size_t)j * d_in;
        double acc = 0.0;
        if (d_in > 1) {
#pragma unroll 2
            for (int c = 0; c < d_in; ++c) acc += ldg(&O[r + (size_t)c * d_out]) * (pw[c] * xc[c]);
        } else acc = ldg(&O[r]) * xc[0];
        Y[e] = acc;
    }
    wsync();
    for (int j = lane_id(); j < m; j += W2B_WARP) {
        double* yc = Y + (size_t)j * d_out;
        double mx = 0.0;
#pragma unroll 1
        for (int r = 0; r < d_out; ++r) mx = fmax(mx, fabs(yc[r]));
        if (mx > 0.0) {
            double la; const double sc = renorm_pow2(mx, la);
#pragma unroll 1
            for (int r = 0; r < d_out; ++r) yc[r] *= sc;
            lg[j] += shift + la;
        }
    }
    wsync();
    return sst_out;
}

// Close ket K (block sst, d x m, scales klg) with bra B (same block; nullptr = unit vectors e_j at tau = beta) across a
// propagation of length dt:  tr_j = sum_c B[c,j] exp(-(E_c-E_0) dt) K[c,j].  pt receives (log, sign) per outer state,
// (tlog, tsign) the signed log-sum (addLogTr, Trace.F90:298-309).
template <class TablesT>
W2B_HDN void mat_close(const TablesT& T, int sst, double dt, const double* K, const double* klg, const double* B, const double* blg,
                      double* pw, int m, double* pt, double& tlog, double& tsign) {
    const int d = ldg(&T.dim[sst]);
    const double* ev = T.evals + ldg(&T.offset[sst]);
    const double e0 = ldg(&ev[0]);
    for (int c = lane_id(); c < d; c += W2B_WARP) pw[c] = c ? w_exp(-(ldg(&ev[c]) - e0) * dt) : 1.0;
    wsync();
    for (int j = lane_id(); j < m; j += W2B_WARP) {
        double s;
        if (B) { s = 0.0;
#pragma unroll 1
            for (int c = 0; c < d; ++c) s += B[c + (size_t)j * d] * pw[c] * K[c + (size_t)j * d]; }
        else s = pw[j] * K[j + (size_t)j * d];
        if (s != 0.0) { pt[2 * j] = klg[j] + (B ? blg[j] : 0.0) - (e0 - T.egs) * dt + w_log(fabs(s)); pt[2 * j + 1] = s < 0.0 ? -1.0 : 1.0; }
        else { pt[2 * j] = NEG_HUGE; pt[2 * j + 1] = 0.0; }
    }
    wsync();
    if (m == 1) { tlog = pt[0]; tsign = pt[1]; return; }
    double mx = NEG_HUGE;
#pragma unroll 1
    for (int j = 0; j < m; ++j) if (pt[2 * j + 1] != 0.0) mx = fmax(mx, pt[2 * j]);
    if (mx == NEG_HUGE) { tlog = NEG_HUGE; tsign = 0.0; return; }
    double sum = 0.0;
#pragma unroll 1
    for (int j = 0; j < m; ++j) if (pt[2 * j + 1] != 0.0) sum += pt[2 * j + 1] * w_exp(pt[2 * j] - mx);
    if (sum == 0.0) { tlog = NEG_HUGE; tsign = 0.0; return; }
    tlog = mx + w_log(fabs(sum)); tsign = sum < 0.0 ? -1.0 : 1.0;
}

}  // namespace w2b
