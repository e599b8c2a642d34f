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

#if W2B_DEVICE
// exact renormalisation of a column whose maximum is zero, subnormal or huge (rare): full reduction over the 8 lanes of the column
W2B_HDN void renorm_slow(double mx, double& sc, double& la, bool& nonzero) {
    mx = fmax(mx, __shfl_xor_sync(W2B_FULL, mx, 4)); mx = fmax(mx, __shfl_xor_sync(W2B_FULL, mx, 8)); mx = fmax(mx, __shfl_xor_sync(W2B_FULL, mx, 16));
    sc = 1.0; la = 0.0; nonzero = mx > 0.0;
    if (nonzero) sc = renorm_pow2(mx, la);
}
#endif

// one packed record per (superstate, operator): everything an operator product needs, fetched with two 16-byte loads
// op_off: the block as a compact column-major d_out x d_in matrix; frag_off: the same block zero-padded and laid out in DMMA
// A-fragment order [k-step][row-tile pair][lane][2] (Tables::opfrag), one coalesced 16-byte load per lane and tile pair
struct alignas(16) OpDesc { int sst_out, d_in, d_out, ev_off; int op_off, frag_off; double e0; };

// X(d_in x m) <- O(d_out x d_in) * diag(exp(-(E_c - E_0) dt)) * X, in place; every column is renormalised by a power
// of two near its largest element and lg[j] += -(E_0 - Egs) dt + e ln2.  `pw` is warp scratch for the d_in propagation
// factors.  Returns the target block.
// Latency / code-size structure (the chains are latency- and instruction-fetch-bound, not throughput-bound): the
// descriptor is one load; the k-loop is rolled (small code) and holds all <= 4 x 4 output tiles in registers, so every A
// fragment is fetched once, one k-step ahead of its use, and X is overwritten only after the last k-step.
// slot / slot_lg (optional): the renormalised result and its log scales are also written to this cache slot (compact d_out x m)
template <int NTMAX, class TablesT>
W2B_HD int mat_apply_inl(const TablesT& T, int sst_in, int opc, double dt, double* X, int ld, double* lg, double* pw, int m,
                         double* slot = nullptr, double* slot_lg = nullptr, int slot_cols = 0, int* work = nullptr) {
#if W2B_DEVICE
    const int4* dp = (const int4*)&T.desc[sst_in * T.nops + opc];
    const int4 q0 = __ldg(dp), q1 = __ldg(dp + 1);
    const int sst_out = q0.x, d_in = q0.y, d_out = q0.z;
    const double* ev = T.evals + q0.w;
    const double* O = T.opdata + q1.x;
    const double2* F = (const double2*)(T.opfrag + q1.y);
    const double e0 = __hiloint2double(q1.w, q1.z);
#else
    const OpDesc& D = T.desc[sst_in * T.nops + opc];
    const int sst_out = D.sst_out, d_in = D.d_in, d_out = D.d_out;
    const double* ev = T.evals + D.ev_off;
    const double* O = T.opdata + D.op_off;
    const double e0 = D.e0;
#endif
    const double shift = -(e0 - T.egs) * dt;
    if (slot) slot += slot_cols * d_out;                          // first of this call's columns inside the cache slot
    if (work) *work += d_in * d_out;
    if (d_in == 1 && d_out == 1) {                               // scalar link of the chain: one multiply per outer state
        const double o = ldg(&O[0]);
        for (int j = lane_id(); j < m; j += W2B_WARP) {
            const double y = o * X[(size_t)j * ld];
            double v = 0.0;
            if (y != 0.0) { double la; const double sc = renorm_pow2(fabs(y), la); v = y * sc; lg[j] += shift + la; }
            X[(size_t)j * ld] = v;
            if (slot) { slot[j] = v; slot_lg[j] = lg[j]; }
        }
        wsync();
        return sst_out;
    }
#if W2B_DEVICE
    const int lane = lane_id(), g = lane >> 2, t = lane & 3;
    const int ntiles = (m + 7) >> 3, mtiles = (d_out + 7) >> 3, ksteps = (d_in + 3) >> 2;
    // A fragments of the first k-step are requested before the exp() of the propagation factors
    const int npair = (mtiles + 1) >> 1;
    double an[4];
    { const double2 v = __ldg(F + lane); an[0] = v.x; an[1] = v.y; }
    if (npair > 1) { const double2 v = __ldg(F + 32 + lane); an[2] = v.x; an[3] = v.y; } else { an[2] = 0.0; an[3] = 0.0; }
    for (int c = lane; c < d_in; c += W2B_WARP) pw[c] = c ? w_exp(-(ldg(&ev[c]) - e0) * dt) : 1.0;
    wsync();
    double acc[NTMAX][4][2];                                       // [column tile][row tile][fragment half]
#pragma unroll
    for (int nt = 0; nt < NTMAX; ++nt)
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) { acc[nt][mt][0] = 0.0; acc[nt][mt][1] = 0.0; }
#pragma unroll 1
    for (int ks = 0; ks < ksteps; ++ks) {
        double ac[4];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) ac[mt] = an[mt];
        if (ks + 1 < ksteps) {                                     // prefetch the next k-step's A fragments
            const double2* f = F + ((ks + 1) * npair) * 32 + lane;
            { const double2 v = __ldg(f); an[0] = v.x; an[1] = v.y; }
            if (npair > 1) { const double2 v = __ldg(f + 32); an[2] = v.x; an[3] = v.y; }
        }
        const int k = ks * 4 + t;
        const bool kok = k < d_in;
        const double pk = kok ? pw[k] : 0.0;
#pragma unroll
        for (int nt = 0; nt < NTMAX; ++nt) {
            if (nt < ntiles) {                                     // warp-uniform
                const int n = nt * 8 + g;
                const double b = (kok && n < m) ? X[n * ld + k] * pk : 0.0;
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
                    if (mt < mtiles) dmma884(acc[nt][mt][0], acc[nt][mt][1], ac[mt], b);
            }
        }
    }
    wsync();                                                       // every lane has consumed X: overwrite in place
#pragma unroll
    for (int nt = 0; nt < NTMAX; ++nt) {
        if (nt < ntiles) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int col = nt * 8 + 2 * t + q;
                double mx = 0.0;
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) mx = fmax(mx, fabs(acc[nt][mt][q]));     // unused row tiles hold zeros
                // column maximum over the 8 lanes sharing t: order-preserving compare of the high words gives the exponent
                int hi = __double2hiint(mx);
                hi = max(hi, __shfl_xor_sync(W2B_FULL, hi, 4)); hi = max(hi, __shfl_xor_sync(W2B_FULL, hi, 8)); hi = max(hi, __shfl_xor_sync(W2B_FULL, hi, 16));
                const int ef = hi >> 20;
                double sc, la; bool nonzero = true;
                if (wany(col < m && (ef == 0 || ef > 2040))) renorm_slow(mx, sc, la, nonzero);   // zero / subnormal / huge column
                else { const int e = ef - 1023; sc = __hiloint2double((1023 - e) << 20, 0); la = (double)e * LN2; }
                if (col < m) {
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) {
                        const int r = mt * 8 + g;
                        if (mt < mtiles && r < d_out) { const double v = acc[nt][mt][q] * sc; X[col * ld + r] = v; if (slot) slot[col * d_out + r] = v; }
                    }
                    if (g == 0) { double l = lg[col]; if (nonzero) { l += shift + la; lg[col] = l; } if (slot) slot_lg[col] = l; }
                }
            }
        }
    }
    wsync();
#else
    for (int c = 0; c < d_in; ++c) pw[c] = c ? w_exp(-(ev[c] - e0) * dt) : 1.0;
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
        for (int r = 0; r < d_out; ++r) { xc[r] = y[r] * sc; if (slot) slot[(size_t)j * d_out + r] = xc[r]; }
        if (slot) slot_lg[j] = lg[j];
    }
#endif
    return sst_out;
}

template <class TablesT>
W2B_HDN int mat_apply(const TablesT& T, int sst_in, int opc, double dt, double* X, int ld, double* lg, double* pw, int m) {
    return mat_apply_inl<4>(T, sst_in, opc, dt, X, ld, lg, pw, m);
}

// Close ket K (block sst, d x m, leading dimension ld, scales klg) with bra B (same block, compact d x m; nullptr = unit
// vectors e_{j0+j} at tau = beta) across a propagation of length dt:  tr_j = sum_c B[c,j] exp(-(E_c-E_0) dt) K[c,j].
// pt receives (log, sign) per outer state.  K, klg, B, blg, pt address the first of the m columns handled here; j0 is that
// column's index inside the outer block (the trace kernel works on 8-column tiles).
template <class TablesT>
W2B_HD void mat_close_parts(const TablesT& T, int sst, double dt, const double* K, int ld, const double* klg, const double* B, const double* blg,
                            double* pw, int m, int j0, double* pt) {
#if W2B_DEVICE
    const int4* dp = (const int4*)&T.desc[sst * T.nops];
    const int4 q0 = __ldg(dp), q1 = __ldg(dp + 1);
    const int d = q0.y;
    const double* ev = T.evals + q0.w;
    const double e0 = __hiloint2double(q1.w, q1.z);
#else
    const int d = T.desc[sst * T.nops].d_in;
    const double* ev = T.evals + T.desc[sst * T.nops].ev_off;
    const double e0 = T.desc[sst * T.nops].e0;
#endif
    for (int c = lane_id(); c < d; c += W2B_WARP) pw[c] = c ? w_exp(-(ldg(&ev[c]) - e0) * dt) : 1.0;
    wsync();
    for (int j = lane_id(); j < m; j += W2B_WARP) {
        double s;
        if (B) { s = 0.0;
#pragma unroll 1
            for (int c = 0; c < d; ++c) s += B[c + j * d] * pw[c] * K[c + j * ld]; }
        else s = pw[j0 + j] * K[j0 + j + j * ld];
        if (s != 0.0) { pt[2 * j] = klg[j] + (B ? blg[j] : 0.0) - (e0 - T.egs) * dt + w_log(fabs(s)); pt[2 * j + 1] = s < 0.0 ? -1.0 : 1.0; }
        else { pt[2 * j] = NEG_HUGE; pt[2 * j + 1] = 0.0; }
    }
    wsync();
}
// signed log-sum of the per-outer-state parts (addLogTr, Trace.F90:298-309), lane-parallel
W2B_HDN void parts_sum(const double* pt, int m, double& tlog, double& tsign) {
    if (m == 1) { tlog = pt[0]; tsign = pt[1]; return; }
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
template <class TablesT>
W2B_HDN void mat_close(const TablesT& T, int sst, double dt, const double* K, int ld, const double* klg, const double* B, const double* blg,
                      double* pw, int m, double* pt, double& tlog, double& tsign) {
    mat_close_parts(T, sst, dt, K, ld, klg, B, blg, pw, m, 0, pt);
    parts_sum(pt, m, tlog, tsign);
}

}  // namespace w2b
