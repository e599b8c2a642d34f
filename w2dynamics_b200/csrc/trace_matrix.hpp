// Blocked local trace in the H_loc eigenbasis (quantum-number blocks larger than 1x1: configs C3, C4).
//
// Reference semantics (files under /root/reference/src/ctqmc_fortran):
//   get_Trace_EB / get_Trace_EB_fullmatrix   Trace.F90:5354-5549, 5734-6033  (ket/bra chains, split_log renormalisation)
//   update_trace_EB                          Trace.F90:6109-6209              (cache hand-over on accept)
//   update_b2_states_eb                      Trace.F90:6213-6469              (states at beta/2 for the density matrix)
//
// B200-first design (not the reference's per-operator `state/cache` pointers tied to the sliding window):
//   For the outer superstate of dimension m the trace is sum_j <j| P O_n P ... O_1 P |j>.  Every operator owns one cache
//   slot holding KET_i = O_i P ... O_1 P I_m (d_i x m, one log scale per column) and BRA_i = (I_m P O_n ... P O_i)^T
//   (d_{i-1} x m).  Two frontiers say which are valid: kets for indices < kv, bras for indices >= bv.  A proposal touching
//   positions [i1, i2] extends the frontiers lazily to i1-1 / i2 (work that is kept whether or not the move is accepted),
//   multiplies through the changed sub-string only and closes with one dot product per outer state.  An accepted move just
//   moves the frontiers (kv = i1, bv = i2 +- 2); nothing is copied.  This is exact (same products as a from-scratch
//   evaluation) and independent of the sliding window.  Slots are addressed through a per-operator slot id so that
//   inserting / removing operators shifts 2-byte ids, not matrices.  Slot storage is compact (leading dimension = actual
//   block dimension).
//
//   The product itself, Y = O * diag(exp(-(E-E0) dt)) * X, runs on the FP64 tensor pipe in tiles of 8 columns (outer
//   states): columns never mix, so a tile is an independent unit of work.  One warp owns one tile in shared memory
//   (column-major, leading dimension ld == 4 mod 16 doubles: conflict-free B-fragment loads) and issues
//   mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4 -- sm_100a lowers every f64 mma shape to it, tools/probes/dmma_probe.cu; tcgen05
//   has no FP64 kind).  What the profile of round 1 showed (profiles/r2b_*): 99 issued instructions per DMMA, two thirds of
//   them integer / address / control, 76 per k-step.  Hence the shape of the code below:
//     * two instances of the product (up to 16 / up to 32 output rows): no predicated DMMA, all accumulators and the A
//       fragments of the current and the next k-step in registers; the whole kernel is kept near 2 k SASS instructions because
//       the warps of an SM sit at unrelated program counters and share a 32 KB instruction cache;
//     * A fragments come from a host-prepared array in fragment order (one coalesced 16-byte load per lane and tile pair);
//     * propagation factors are zero-padded to the k range, the tile is zero-padded to the tile shape: no bounds tests in
//       the loop;
//     * ONE power-of-two scale per tile and step (integer max over the exponent fields + 5 shuffles) instead of one per
//       column: scaling by 2^-e is exact, so any scale that keeps the numbers in range gives bit-identical products; the
//       per-column log scales all receive the same e ln2.  (The reference divides by the largest element of each vector
//       and adds its log, Trace.F90:5802-5822 / split_log 5341-5348.)
#pragma once
#include "warp.hpp"

namespace w2b {

struct Tables;
struct Chains;
struct Ctx;

constexpr double NEG_HUGE = -1.7976931348623157e308;
constexpr int MAT_DMAX = 32;          // largest block dimension of the tensor path (4 row tiles in registers)
constexpr int TILE_COLS = 8;          // outer states (columns) per tile = N of the DMMA shape

// leading dimension of a work matrix for blocks up to dmax: the smallest value >= 8 ceil(dmax / 8) that is 4 mod 16
#if W2B_DEVICE
__host__ __device__
#endif
inline int mat_ld(int dmax) { const int need = (dmax + 7) & ~7; int ld = 4; while (ld < need) ld += 16; return ld; }
// columns a work matrix must provide for an outer block of dimension dmax (whole tiles)
#if W2B_DEVICE
__host__ __device__
#endif
inline int mat_cols(int dmax) { return (dmax + 7) & ~7; }

// X = I_m (columns up to the next tile boundary zero), log scales 0
W2B_HDN void mat_identity(double* X, int ld, double* lg, int m) {
    const int mc = W2B_DEVICE ? mat_cols(m) : m;
    for (int j = 0; j < mc; ++j) for (int r = lane_id(); r < ld && r < mat_cols(m); r += W2B_WARP) X[r + j * ld] = (r == j && j < m) ? 1.0 : 0.0;
    for (int j = lane_id(); j < mc; j += W2B_WARP) lg[j] = 0.0;
    wsync();
}
// compact m x m identity (leading dimension m), no tile padding
W2B_HDN void mat_identity_compact(double* X, double* lg, int m) {
    for (int e = lane_id(); e < m * m; e += W2B_WARP) X[e] = (e % m == e / m) ? 1.0 : 0.0;
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
// load a compact d x m block into a work matrix and zero the padding of its tiles (rows d .. 8 ceil(d/8), columns m .. 8 ceil(m/8))
W2B_HDN void mat_load_padded(double* X, int ld, const double* src, int d, int m) {
#if W2B_DEVICE
    const int R = mat_cols(d), mc = mat_cols(m);
    for (int j = 0; j < mc; ++j) for (int r = lane_id(); r < R; r += W2B_WARP) X[r + (size_t)j * ld] = (r < d && j < m) ? src[r + (size_t)j * d] : 0.0;
#else
    for (int j = 0; j < m; ++j) for (int r = 0; r < d; ++r) X[r + (size_t)j * ld] = src[r + (size_t)j * d];
#endif
}

// power-of-two renormalisation: returns the scale 2^-e, e = exponent of mx; lgadd = e ln 2
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

// one packed record per (superstate, operator): everything an operator product needs, fetched with two 16-byte loads
// op_off: the block as a compact column-major d_out x d_in matrix; frag_off: the same block zero-padded and laid out in DMMA
// A-fragment order [k-step][row-tile pair][lane][2] (Tables::opfrag), one coalesced 16-byte load per lane and tile pair
struct alignas(16) OpDesc { int sst_out, d_in, d_out, ev_off; int op_off, frag_off; double e0; };

#if W2B_DEVICE
// ------------------------------------------------------------------ device: one warp, one 8-column tile
// work space of a tile: X (ld x 8, column-major), lg[8] (log scale per column), pw[32] (propagation factors)
struct Tile { double* X; double* lg; double* pw; int ld; };

// D(8x8) += A(8x4, row) * B(4x8, col).  Fragments: lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2*(l%4) + {0,1}].
W2B_HD void dmma884(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
W2B_HD int expfield(double x) { return __double2hiint(x) & 0x7ff00000; }
// tile-uniform scale from the largest exponent field e (0 = all zero / subnormal: leave as is)
W2B_HD void tile_scale(int e, double& sc, double& la) {
    sc = 1.0; la = 0.0;
    if (e != 0) { sc = __hiloint2double(0x7fe00000 - e, 0); la = (double)((e >> 20) - 1023) * LN2; }
}

// acc(8 MT x 8) = O(8 MT x 4 ksteps) * (pw o X): the k-loop.  xb / pwb point at this lane's first B element / factor.
// acc(8 MT x 8) = O(8 MT x 4 ksteps) * (pw o X): the k-loop.  xb / pwb point at this lane's first B element / factor; np = row-tile
// pairs the fragment array holds per k-step (1 or 2).  Two instances only (MT = 2: blocks up to 16 rows, MT = 4: up to 32): the
// kernel's hot code has to stay inside the 32 KB instruction cache of the SM -- with one instance per row-tile count (1..4) and
// a two-fold unrolled loop the kernel was 3960 SASS instructions and 25 % slower than this 2270-instruction version although
// it executed fewer instructions (profiles/r2g_*).
template <int MT>
W2B_HD void tile_mma(const double2* __restrict__ F, int ksteps, const double* xb, const double* pwb, double (&acc)[MT][2], int np) {
    double2 a0 = __ldg(F), a1 = make_double2(0.0, 0.0), n0 = a0, n1 = a1;
    if (MT > 2 && np > 1) { a1 = __ldg(F + 32); n1 = a1; }
#pragma unroll 1
    for (int ks = 0; ks < ksteps; ++ks) {
        F += np * 32;
        if (ks + 1 < ksteps) { n0 = __ldg(F); if (MT > 2 && np > 1) n1 = __ldg(F + 32); }      // A fragments of the next k-step
        const double b = xb[4 * ks] * pwb[4 * ks];
        dmma884(acc[0][0], acc[0][1], a0.x, b); dmma884(acc[1][0], acc[1][1], a0.y, b);
        if (MT > 2) { dmma884(acc[2][0], acc[2][1], a1.x, b); dmma884(acc[3][0], acc[3][1], a1.y, b); }
        a0 = n0; a1 = n1;
    }
}

template <int MT>
W2B_HD void tile_apply_mt(const double2* __restrict__ F, int ksteps, int d_out, double shift, const Tile& w, int mw, double* slot, double* slot_lg) {
    const int lane = lane_id(), g = lane >> 2, t = lane & 3;
    double acc[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) { acc[mt][0] = 0.0; acc[mt][1] = 0.0; }
    tile_mma<MT>(F + lane, ksteps, w.X + g * w.ld + t, w.pw + t, acc, (((d_out + 7) >> 3) + 1) >> 1);
    int e = 0;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) e = max(e, max(expfield(acc[mt][0]), expfield(acc[mt][1])));
    e = wmax(e);
    double sc, la;
    tile_scale(e, sc, la);
    wsync();                                                     // every lane has consumed X: overwrite in place
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int col = 2 * t + q;
        double* xc = w.X + col * w.ld + g;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) xc[8 * mt] = acc[mt][q] * sc;      // all rows of the MT tiles: the padding rows become zero
        if (slot && col < mw) {
            double* sc_ = slot + col * d_out + g;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) if (8 * mt + g < d_out) sc_[8 * mt] = acc[mt][q] * sc;
        }
    }
    if (lane < TILE_COLS) { const double l = w.lg[lane] + shift + la; w.lg[lane] = l; if (slot && lane < mw) slot_lg[lane] = l; }
    wsync();
}

// X(d_in x 8) <- O(d_out x d_in) * diag(exp(-(E_c - E_0) dt)) * X in place, tile renormalised by a power of two,
// lg[j] += -(E_0 - Egs) dt + e ln2.  slot / slot_lg (optional): the result and its log scales also go to this cache slot
// (compact d_out x m; slot points at the slot's first column, j0 = first column of this tile, slot_lg at its first scale).
// mw: valid columns of the tile.  Returns the target block.
template <class TablesT>
W2B_HD int tile_apply(const TablesT& T, int sst_in, int opc, double dt, const Tile& w, int mw, double* slot, double* slot_lg, int j0, int* work) {
    const int4* dp = (const int4*)&T.desc[sst_in * T.nops + opc];
    const int4 q0 = __ldg(dp), q1 = __ldg(dp + 1);
    const int sst_out = q0.x, d_in = q0.y, d_out = q0.z;
    const double e0 = __hiloint2double(q1.w, q1.z);
    const double shift = -(e0 - T.egs) * dt;
    const int lane = lane_id();
    if (slot) slot += j0 * d_out;
    if (work) *work += d_in * d_out;
    if (d_in == 1 && d_out == 1) {                               // scalar link of the chain: one multiply per outer state
        const double o = __ldg(T.opdata + q1.x);
        const int col = lane >> 2, r = lane & 3;
        const double y = r == 0 ? o * w.X[col * w.ld] : 0.0;
        const int e = wmax(expfield(y));
        double sc, la;
        tile_scale(e, sc, la);
        wsync();
        w.X[col * w.ld + r] = y * sc;                            // rows 1..3 zero: k padding of the next product
        if (slot && r == 0 && col < mw) slot[col] = y * sc;
        if (lane < TILE_COLS) { const double l = w.lg[lane] + shift + la; w.lg[lane] = l; if (slot && lane < mw) slot_lg[lane] = l; }
        wsync();
        return sst_out;
    }
    {   // propagation factors, zero beyond d_in (k padding)
        double p = 0.0;
        if (lane < d_in) p = lane ? w_exp(-(__ldg(T.evals + q0.w + lane) - e0) * dt) : 1.0;
        w.pw[lane] = p;
    }
    wsync();
    const double2* F = (const double2*)(T.opfrag + q1.y);
    const int ksteps = (d_in + 3) >> 2;
    if (d_out <= 16) tile_apply_mt<2>(F, ksteps, d_out, shift, w, mw, slot, slot_lg);
    else tile_apply_mt<4>(F, ksteps, d_out, shift, w, mw, slot, slot_lg);
    return sst_out;
}

// tile <- columns j0 .. j0+mw-1 of a cached d x m block (src points at column j0, src_lg at its scale) or, src == nullptr,
// of the identity I_m; padding rows / columns of the tile are zeroed
W2B_HDN void tile_load(const Tile& w, const double* src, const double* src_lg, int d, int j0, int mw) {
    const int lane = lane_id(), R = mat_cols(d);
#pragma unroll 1
    for (int col = 0; col < TILE_COLS; ++col) {
        double v = 0.0;
        if (src) { if (lane < d && col < mw) v = src[col * d + lane]; }
        else if (col < mw && lane == j0 + col) v = 1.0;
        if (lane < R) w.X[col * w.ld + lane] = v;
    }
    if (lane < TILE_COLS) w.lg[lane] = (src && lane < mw) ? src_lg[lane] : 0.0;
    wsync();
}

// Close the ket tile (block sst, d x 8) with bra columns B (same block, compact d x m, pointing at column j0; nullptr = unit
// vectors e_{j0+j} at tau = beta) across a propagation of length dt:  tr_j = sum_c B[c,j] exp(-(E_c-E_0) dt) K[c,j].
// pt receives (log, sign) per outer state of the tile.
template <class TablesT>
W2B_HD void tile_close(const TablesT& T, int sst, double dt, const Tile& w, const double* B, const double* blg, int mw, int j0, double* pt) {
    const int4* dp = (const int4*)&T.desc[sst * T.nops];
    const int4 q0 = __ldg(dp), q1 = __ldg(dp + 1);
    const int d = q0.y, lane = lane_id(), g = lane >> 2, t = lane & 3;
    const double e0 = __hiloint2double(q1.w, q1.z);
    {
        double p = 0.0;
        if (lane < d) p = lane ? w_exp(-(__ldg(T.evals + q0.w + lane) - e0) * dt) : 1.0;
        w.pw[lane] = p;
    }
    wsync();
    double s = 0.0;
    if (B) {
        if (g < mw) for (int c = t; c < d; c += 4) s += B[c + g * d] * w.pw[c] * w.X[g * w.ld + c];
        s += __shfl_xor_sync(W2B_FULL, s, 1); s += __shfl_xor_sync(W2B_FULL, s, 2);
    } else if (g < mw) s = w.pw[j0 + g] * w.X[g * w.ld + j0 + g];
    if (t == 0 && g < mw) {
        if (s != 0.0) { pt[2 * g] = w.lg[g] + (B ? blg[g] : 0.0) - (e0 - T.egs) * dt + w_log(fabs(s)); pt[2 * g + 1] = s < 0.0 ? -1.0 : 1.0; }
        else { pt[2 * g] = NEG_HUGE; pt[2 * g + 1] = 0.0; }
    }
    wsync();
}
#endif  // W2B_DEVICE

// ------------------------------------------------------------------ whole-matrix interface (all m columns of an outer block)
// Used by the paths that are not the trace kernel: set_mc_config rebuilds, the measurement's fallback frontier extension, the
// host-emulation build.  On the device it runs the tile code above over the column tiles of X (ld x 8 ceil(m/8), padding
// zeroed by mat_identity / mat_load_padded); on the host a plain loop nest with per-column renormalisation.
// X(d_in x m) <- O(d_out x d_in) * diag(exp(-(E_c - E_0) dt)) * X, in place; lg[j] += -(E_0 - Egs) dt + scale.  `pw`: 32 doubles of
// warp scratch.  slot / slot_lg (optional): the result also goes to this cache slot (compact d_out x m, columns from slot_cols on).
template <class TablesT>
W2B_HD int mat_apply_slot(const TablesT& T, int sst_in, int opc, double dt, double* X, int ld, double* lg, double* pw, int m,
                          double* slot, double* slot_lg, int slot_cols, int* work) {
#if W2B_DEVICE
    int out = -1;
    for (int j0 = 0; j0 < m; j0 += TILE_COLS) {
        const Tile w{X + (size_t)j0 * ld, lg + j0, pw, ld};
        const int mw = m - j0 < TILE_COLS ? m - j0 : TILE_COLS;
        out = tile_apply(T, sst_in, opc, dt, w, mw, slot, slot ? slot_lg + j0 : nullptr, slot_cols + j0, j0 == 0 ? work : nullptr);
    }
    return out;
#else
    const OpDesc& D = T.desc[sst_in * T.nops + opc];
    const int sst_out = D.sst_out, d_in = D.d_in, d_out = D.d_out;
    const double* ev = T.evals + D.ev_off;
    const double* O = T.opdata + D.op_off;
    const double e0 = D.e0;
    const double shift = -(e0 - T.egs) * dt;
    if (slot) slot += slot_cols * d_out;
    if (work) *work += d_in * d_out;
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
    return sst_out;
#endif
}
template <class TablesT>
W2B_HDN int mat_apply(const TablesT& T, int sst_in, int opc, double dt, double* X, int ld, double* lg, double* pw, int m) {
    return mat_apply_slot(T, sst_in, opc, dt, X, ld, lg, pw, m, (double*)nullptr, (double*)nullptr, 0, (int*)nullptr);
}

// Close ket K (block sst, d x m, leading dimension ld, scales klg) with bra B (same block, compact d x m; nullptr = unit
// vectors e_{j0+j} at tau = beta) across a propagation of length dt:  tr_j = sum_c B[c,j] exp(-(E_c-E_0) dt) K[c,j].
// pt receives (log, sign) per outer state.  K, klg, B, blg, pt address the first of the m columns handled here; j0 is that
// column's index inside the outer block.
template <class TablesT>
W2B_HD void mat_close_parts(const TablesT& T, int sst, double dt, double* K, int ld, double* klg, const double* B, const double* blg,
                            double* pw, int m, int j0, double* pt) {
#if W2B_DEVICE
    const int d = ldg(&T.dim[sst]);
    for (int c0 = 0; c0 < m; c0 += TILE_COLS) {
        const Tile w{K + (size_t)c0 * ld, klg + c0, pw, ld};
        const int mw = m - c0 < TILE_COLS ? m - c0 : TILE_COLS;
        tile_close(T, sst, dt, w, B ? B + (size_t)c0 * d : nullptr, B ? blg + c0 : nullptr, mw, j0 + c0, pt + 2 * c0);
    }
#else
    const int d = T.desc[sst * T.nops].d_in;
    const double* ev = T.evals + T.desc[sst * T.nops].ev_off;
    const double e0 = T.desc[sst * T.nops].e0;
    for (int c = 0; c < d; ++c) pw[c] = c ? w_exp(-(ev[c] - e0) * dt) : 1.0;
    for (int j = 0; j < m; ++j) {
        double s;
        if (B) { s = 0.0; for (int c = 0; c < d; ++c) s += B[c + j * d] * pw[c] * K[c + j * ld]; }
        else s = pw[j0 + j] * K[j0 + j + j * ld];
        if (s != 0.0) { pt[2 * j] = klg[j] + (B ? blg[j] : 0.0) - (e0 - T.egs) * dt + w_log(std::fabs(s)); pt[2 * j + 1] = s < 0.0 ? -1.0 : 1.0; }
        else { pt[2 * j] = NEG_HUGE; pt[2 * j + 1] = 0.0; }
    }
#endif
}
// signed log-sum of the per-outer-state parts (addLogTr, Trace.F90:298-309), lane-parallel
// only >= 0 (state sampling, Trace.F90:5759-5764): the contribution of that single outer state is the trace
W2B_HDN void parts_sum(const double* pt, int m, double& tlog, double& tsign, int only = -1) {
    if (only >= 0) { tsign = pt[2 * only + 1]; tlog = tsign != 0.0 ? pt[2 * only] : NEG_HUGE; return; }
    if (m == 1) { tlog = pt[0]; tsign = pt[1]; return; }
    double mx = NEG_HUGE;
#pragma unroll 1
    for (int j = lane_id(); j < m; j += W2B_WARP) if (pt[2 * j + 1] != 0.0) mx = fmax(mx, pt[2 * j]);
    mx = wmaxd(mx);
    if (mx == NEG_HUGE) { tlog = NEG_HUGE; tsign = 0.0; return; }
    double sum = 0.0;
#pragma unroll 1
    for (int j = lane_id(); j < m; j += W2B_WARP) if (pt[2 * j + 1] != 0.0) sum += pt[2 * j + 1] * w_exp(pt[2 * j] - mx);
    sum = wsumd(sum);
    if (sum == 0.0) { tlog = NEG_HUGE; tsign = 0.0; return; }
    tlog = mx + w_log(fabs(sum)); tsign = sum < 0.0 ? -1.0 : 1.0;
}
template <class TablesT>
W2B_HDN void mat_close(const TablesT& T, int sst, double dt, double* K, int ld, double* klg, const double* B, const double* blg,
                      double* pw, int m, double* pt, double& tlog, double& tsign, int only = -1) {
    mat_close_parts(T, sst, dt, K, ld, klg, B, blg, pw, m, 0, pt);
    parts_sum(pt, m, tlog, tsign, only);
}

}  // namespace w2b
