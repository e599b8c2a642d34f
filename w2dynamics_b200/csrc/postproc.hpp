// Normalisation of the accumulators and the density-matrix estimators, i.e. the part of
// ctqmc_measure that runs once after sampling (reference CTQMC.F90:8663-8771) plus
// Meas_rho1/rho2/DoubleOcc_from_Densitymatrix (2884-3060) and Transform_DensityMatrix_EB_to_OCCB (3163-3179).
// Host code (runs once per solve on a few hundred kB); results are stored column-major with the
// extents of the reference's module arrays.
#pragma once
#include <cmath>
#include <limits>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "atomic.hpp"
#include "engine_core.hpp"

namespace w2b {

struct RawBlock {                 // raw sums of one bin (or of all bins)
    std::vector<double> v;        // accumulator slab (layout of Accum)
};

struct ResultSet {
    std::map<std::string, std::vector<double>> arr;
    std::map<std::string, double> scal;
};

namespace detail {
// C(m x n) = A(m x k) * B(k x n), all column-major
inline std::vector<double> mm(const double* A, int m, int k, const double* B, int n) {
    std::vector<double> Cm((size_t)m * n, 0.0);
    for (int j = 0; j < n; ++j) for (int p = 0; p < k; ++p) { double b = B[p + (size_t)j * k]; if (b == 0.0) continue; for (int i = 0; i < m; ++i) Cm[i + (size_t)j * m] += A[i + (size_t)p * m] * b; }
    return Cm;
}
}  // namespace detail

// Configuration-independent half of Meas_rho2_from_Densitymatrix (CTQMC.F90:2945-3017): for every block k and every pair of
// annihilators (f1 first, then f2) with k -c1-> n1 -c2-> n2 the product  P21 = c2(n1 -> n2) c1(k -> n1)  (d_n2 x d_k).
// rho2(f4,f3,f2,f1) = sum_k Tr[rho_k c4^+ c3^+ P21] then costs one dot product per element instead of two more matrix products.
// Built once per solve (a few MB for the 5-orbital problem).
struct Rho2Cache {
    struct Entry { int f1, f2, n2; std::vector<double> p21; };
    std::vector<std::vector<Entry>> by_block;
    void build(const AtomicTables& at) {
        const int nf = at.nf;
        by_block.assign(at.nsst, {});
        for (int k = 0; k < at.nsst; ++k) for (int f1 = 0; f1 < nf; ++f1) {
            const int n1 = at.conn(k, f1 * 2 + 1);
            if (n1 < 0) continue;
            for (int f2 = 0; f2 < nf; ++f2) {
                const int n2 = at.conn(n1, f2 * 2 + 1);
                if (n2 < 0) continue;
                by_block[k].push_back(Entry{f1, f2, n2, detail::mm(at.op(n1, f2 * 2 + 1), at.dim[n2], at.dim[n1], at.op(k, f1 * 2 + 1), at.dim[k])});
            }
        }
    }
};

// Normalise one slab.  dens_eb: block-diagonal eigenbasis density matrix, block k column-major at dens_off[k].
inline void normalise_slab(const AtomicTables& at, const Accum& A, const Run& R, double beta, const double* slab,
                           const std::vector<int>& dens_off, bool want_rho2, ResultSet& out, const Rho2Cache* r2c = nullptr) {
    const int nb = at.nb, nf = at.nf, ntau = R.ntau, nleg = R.nlegmax, nh = R.maxhisto + 1, ns = at.nstates, nsst = at.nsst;
    const double cnt = slab[A.o_cnt];
    const double mean_sign = slab[A.o_sign] / cnt;                           // CTQMC.F90:8670
    const double zs = -1.0 / (cnt * mean_sign);                              // zScaling 8697
    out.scal["mean_sign"] = mean_sign;
    out.scal["cnt_meas"] = cnt;
    auto& gtau = out.arr["gtau"]; gtau.assign((size_t)nf * ntau, 0.0);
    auto& gleg = out.arr["gleg"]; gleg.assign((size_t)nf * nleg, 0.0);
    auto& histo = out.arr["histo"]; histo.assign((size_t)nf * nh, 0.0);
    for (int b = 0; b < nb; ++b) for (int s = 0; s < 2; ++s) {
        int f = 2 * b + s; size_t bs = b + (size_t)nb * s;
        for (int t = 0; t < ntau; ++t)                                         // Gtau = -mean/(-mean_sign) * Ntau/beta^2  (8730)
            gtau[bs + (size_t)nf * t] = -(slab[A.o_gtau + f * ntau + t] / cnt) / (-mean_sign) * (double)ntau / (beta * beta);
        for (int l = 0; l < nleg; ++l)                                         // GLeg(l) *= sqrt(2l+1)/beta * (-zScaling) (8766-8770)
            gleg[bs + (size_t)nf * l] = -(slab[A.o_gleg + f * nleg + l] * std::sqrt(2.0 * l + 1.0) / beta) * zs;
        for (int k = 0; k < nh; ++k) histo[bs + (size_t)nf * k] = slab[A.o_histo + f * nh + k] / cnt;   // 8755
    }
    if (R.ngiw > 0) {                                                          // Giw *= zScaling / beta (CTQMC.F90:8815-8817); (re, im) pairs, element (b, s, n) Fortran order
        const int nfreq = 2 * R.ngiw;
        auto& giw = out.arr["giw"]; giw.assign((size_t)2 * nf * nfreq, 0.0);
        for (int b = 0; b < nb; ++b) for (int s = 0; s < 2; ++s) for (int m = 0; m < nfreq; ++m) {
            const size_t e = (b + (size_t)nb * s) + (size_t)nf * m, src = (size_t)A.o_giw + 2 * ((size_t)(2 * b + s) * nfreq + m);
            giw[2 * e] = slab[src] * zs / beta; giw[2 * e + 1] = slab[src + 1] * zs / beta;
        }
    }
    auto div = [](double a, double b) { return a / b; };
    auto& oh = out.arr["outersuperstatehisto"]; oh.resize(nsst); auto& sh = out.arr["signhistosuperstates"]; sh.resize(nsst);
    auto& ch = out.arr["tracecontribsuperstates"]; ch.resize(nsst);
    for (int k = 0; k < nsst; ++k) { sh[k] = div(slab[A.o_signsst + k], slab[A.o_outer + k]); oh[k] = slab[A.o_outer + k] / cnt; ch[k] = slab[A.o_contribsst + k] / cnt; }
    auto& os = out.arr["outerstatehisto"]; os.resize(ns); auto& ss = out.arr["signhistostates"]; ss.resize(ns);
    auto& cs = out.arr["tracecontribstates"]; cs.resize(ns);
    for (int k = 0; k < ns; ++k) { ss[k] = div(slab[A.o_signst + k], slab[A.o_outerst + k]); os[k] = slab[A.o_outerst + k] / cnt; cs[k] = slab[A.o_contribst + k] / cnt; }

    // ---- density matrix estimators (8701-8717)
    std::vector<std::vector<double>> rho(nsst);
    for (int k = 0; k < nsst; ++k) { int d = at.dim[k]; rho[k].resize((size_t)d * d); for (int i = 0; i < d * d; ++i) rho[k][i] = slab[A.o_dens + dens_off[k] + i] * zs; }
    auto& occ_out = out.arr["occ"]; occ_out.assign((size_t)nf * nf, 0.0);
    auto& rho1_out = out.arr["rho1"]; rho1_out.assign((size_t)nf * nf, 0.0);
    auto& rho2_out = out.arr["rho2"]; rho2_out.assign((size_t)nf * nf * nf * nf, 0.0);
    auto opc = [&](int b, int s, int ca) { return (b * 2 + s) * 2 + (ca - 1); };
    auto trace_with = [&](const std::vector<double>& X, int rows, int cols, const double* B) {   // trace(X(rows x cols) * B(cols x rows))
        double t = 0; for (int i = 0; i < rows; ++i) for (int p = 0; p < cols; ++p) t += X[i + (size_t)p * rows] * B[p + (size_t)i * cols]; return t; };
    // the blocks are independent: host threads take blocks round-robin and sum into private copies (a few hundred small matrix
    // products per block for the 5-orbital problem; single-threaded this was half of an end-to-end solve of one accumulation block)
    int nthr = (int)std::thread::hardware_concurrency(); if (nthr < 1) nthr = 1; if (nthr > 8) nthr = 8; if (nthr > nsst) nthr = nsst;
    std::vector<std::vector<double>> p_occ(nthr, std::vector<double>(occ_out.size(), 0.0)), p_rho1(nthr, std::vector<double>(rho1_out.size(), 0.0)),
                                     p_rho2(nthr, std::vector<double>(want_rho2 ? rho2_out.size() : 0, 0.0));
    auto work = [&](int tid) {
    std::vector<double>& occ = p_occ[tid]; std::vector<double>& rho1 = p_rho1[tid]; std::vector<double>& rho2 = p_rho2[tid];
    for (int k = tid; k < nsst; k += nthr) {
        const int d = at.dim[k];
        for (int b4 = 0; b4 < nb; ++b4) for (int s4 = 0; s4 < 2; ++s4) {
            // t1 = rho * c4^+ (from n3 -> k)  ; requires some n3 with Connect(n3, 4, c^dag) == k  <=> Connect(k, 4, c) == n3
            int n3 = at.conn(k, opc(b4, s4, 2));
            if (n3 < 0) continue;
            int d3 = at.dim[n3];
            std::vector<double> t1 = detail::mm(rho[k].data(), d, d, at.op(n3, opc(b4, s4, 1)), d3);      // d x d3
            // rho1(4, 1) += trace(t1 * c1(k -> n3))  for flavours 1 with Connect(k,1,c) == n3   (Meas_rho1 3020-3060)
            for (int b1 = 0; b1 < nb; ++b1) for (int s1 = 0; s1 < 2; ++s1) {
                if (at.conn(k, opc(b1, s1, 2)) != n3) continue;
                rho1[(2 * b4 + s4) + (size_t)nf * (2 * b1 + s1)] += trace_with(t1, d, d3, at.op(k, opc(b1, s1, 2)));
            }
            // occ(1,2) = Tr[rho c2^+ c2 c1^+ c1]  (MeasDoubleOcc 2884-2942): here flavour 4 plays "2"
            {
                std::vector<double> t2 = detail::mm(t1.data(), d, d3, at.op(k, opc(b4, s4, 2)), d);       // d x d
                for (int b1 = 0; b1 < nb; ++b1) for (int s1 = 0; s1 < 2; ++s1) {
                    int n1 = at.conn(k, opc(b1, s1, 2));
                    if (n1 < 0) continue;
                    int d1 = at.dim[n1];
                    std::vector<double> t3 = detail::mm(t2.data(), d, d, at.op(n1, opc(b1, s1, 1)), d1);  // d x d1
                    size_t idx = b1 + (size_t)nb * (s1 + 2 * (b4 + (size_t)nb * s4));
                    occ[idx] += trace_with(t3, d, d1, at.op(k, opc(b1, s1, 2)));
                }
            }
            if (!want_rho2) continue;
            // rho2(4,3,2,1) = Tr[rho c4^+ c3^+ c2 c1] over chains k -c1-> n1 -c2-> n2 -c3^+-> n3 -c4^+-> k  (Meas_rho2 2945-3017)
            for (int b3 = 0; b3 < nb; ++b3) for (int s3 = 0; s3 < 2; ++s3) {
                int n2 = at.conn(n3, opc(b3, s3, 2));              // Connect(n2, 3, c^dag) == n3  <=> Connect(n3, 3, c) == n2
                if (n2 < 0) continue;
                int d2 = at.dim[n2];
                std::vector<double> u2 = detail::mm(t1.data(), d, d3, at.op(n2, opc(b3, s3, 1)), d2);     // rho c4^+ c3^+ : d x d2
                if (r2c) {                                         // Tr[u2 (c2 c1)] = sum_{i,p} u2(i,p) P21(p,i)
                    for (const Rho2Cache::Entry& en : r2c->by_block[k]) {
                        if (en.n2 != n2) continue;
                        size_t idx = (2 * b4 + s4) + (size_t)nf * ((2 * b3 + s3) + (size_t)nf * (en.f2 + (size_t)nf * en.f1));
                        rho2[idx] += trace_with(u2, d, d2, en.p21.data());
                    }
                    continue;
                }
                for (int b2 = 0; b2 < nb; ++b2) for (int s2 = 0; s2 < 2; ++s2) {
                    int n1 = at.conn(n2, opc(b2, s2, 1));          // Connect(n1, 2, c) == n2  <=> Connect(n2, 2, c^dag) == n1
                    if (n1 < 0) continue;
                    int d1 = at.dim[n1];
                    if (at.conn(n1, opc(b2, s2, 2)) != n2) continue;
                    std::vector<double> u3 = detail::mm(u2.data(), d, d2, at.op(n1, opc(b2, s2, 2)), d1); // d x d1
                    for (int b1 = 0; b1 < nb; ++b1) for (int s1 = 0; s1 < 2; ++s1) {
                        if (at.conn(k, opc(b1, s1, 2)) != n1) continue;
                        size_t idx = (2 * b4 + s4) + (size_t)nf * ((2 * b3 + s3) + (size_t)nf * ((2 * b2 + s2) + (size_t)nf * (2 * b1 + s1)));
                        rho2[idx] += trace_with(u3, d, d1, at.op(k, opc(b1, s1, 2)));
                    }
                }
            }
        }
    }
    };
    if (nthr == 1) work(0);
    else { std::vector<std::thread> th; for (int t = 0; t < nthr; ++t) th.emplace_back(work, t); for (auto& t : th) t.join(); }
    for (int t = 0; t < nthr; ++t) {
        for (size_t i = 0; i < occ_out.size(); ++i) occ_out[i] += p_occ[t][i];
        for (size_t i = 0; i < rho1_out.size(); ++i) rho1_out[i] += p_rho1[t][i];
        for (size_t i = 0; i < p_rho2[t].size(); ++i) rho2_out[i] += p_rho2[t][i];
    }
    // densitymatrix in the occupation basis (3163-3179), stored (NStates x NStates) in block order
    auto& dm = out.arr["densitymatrix"]; dm.assign((size_t)ns * ns, 0.0);
    for (int k = 0; k < nsst; ++k) {
        int d = at.dim[k], o = at.offset[k];
        std::vector<double> t = detail::mm(at.evec[k].data(), d, d, rho[k].data(), d);
        for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) { double a = 0; for (int p = 0; p < d; ++p) a += t[i + (size_t)p * d] * at.evec[k][j + (size_t)p * d]; dm[(o + i) + (size_t)ns * (o + j)] = a; }
    }
}

}  // namespace w2b
