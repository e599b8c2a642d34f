// ORACLE -- test infrastructure, NOT product code.
//
// CPU restatement of the Z-space CT-HYB Markov chain of w2dynamics' matrix solver
// (segment=0; the four pair / global / tau-shift moves, user symmetry moves, outer pair moves, 4-operator moves, state sampling,
// the Matsubara estimator; offdiag=1 is restated for pair, 4-operator, global and tau-shift moves as groundwork for the engine).  Each function
// cites the reference lines (relative to /root/reference/src/ctqmc_fortran) it
// follows.  Representation differs from the reference on purpose (time-sorted
// arrays instead of linked lists, no ket/bra caches: as a checker the local trace is
// always recomputed from tau=0 to beta, which is mathematically identical to the cached
// evaluation of Trace.F90:5354-6033; an opt-in windowed evaluation with cached edge
// states reproduces the reference's cost structure for the timed CPU baseline only);
// the *semantics* -- RNG draw order, proposal
// probabilities, FPos/NPairs/window bookkeeping, signs, measurement formulas --
// are restated verbatim.
//
// Parity status: the reference cannot be built here (no Fortran compiler) and
// its tests hold no golden vectors for this path: "parity
// unpinned" by the reference.  The pins are (i) analytic limits, (ii) from-scratch
// LU vs fast updates, (iii) the reference's C++ RNG (src/mtrng/random.cc compiled
// unchanged into oracle/_ref/libmtrng.so) whose stream this oracle can consume in
// the reference's draw order, (iv) EXACT results of interacting impurity + bath models
// from dense diagonalisation (tests/_exact_anderson.py): G(tau), Legendre coefficients,
// G(iw), <n_i n_j>, rho1 / rho2, with every optional move type, state sampling and
// off-diagonal hybridisation (tests/test_oracle_pins.py, test_matsubara.py, test_oracle_offdiag.py).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm
// may load this library.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <limits>
#include <vector>

namespace {

const double HUGE_D = std::numeric_limits<double>::max();

struct LogVal { double log, sign; };   // TLogTr / TLogDet (Trace.F90:80-82, MatrixUpdate.F90:5-7)
inline LogVal lmul(LogVal a, LogVal b) { return {a.log + b.log, a.sign * b.sign}; }
inline LogVal ldiv(LogVal a, LogVal b) { return {a.log - b.log, a.sign * b.sign}; }
inline LogVal ladd(LogVal a, LogVal b) {           // addLogTr Trace.F90:298-309
    if (a.log >= b.log) return {a.log + std::log(1.0 + a.sign * b.sign * std::exp(b.log - a.log)), a.sign};
    return {b.log + std::log(1.0 + a.sign * b.sign * std::exp(a.log - b.log)), b.sign};
}
inline double lval(LogVal a) { return a.sign * std::exp(a.log); }
inline double sgn(double x) { return std::signbit(x) ? -1.0 : 1.0; }   // Fortran sign(1.0, x)

// ------------------------------------------------------------------ RNG
struct Philox {                       // counter-based stream shared with the CUDA engine
    uint32_t k0, k1; uint64_t ctr; uint32_t buf[4]; uint64_t have;
    static void round10(uint32_t c[4], uint32_t k0, uint32_t k1) {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
        for (int r = 0; r < 10; ++r) {
            uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
            uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
            uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
            c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3; k0 += W0; k1 += W1;
        }
    }
    double next() {
        uint64_t blk = ctr >> 1;
        if (have != blk) {
            buf[0] = (uint32_t)blk; buf[1] = (uint32_t)(blk >> 32); buf[2] = 0; buf[3] = 0;
            round10(buf, k0, k1); have = blk;
        }
        uint32_t a = buf[(ctr & 1) * 2], b = buf[(ctr & 1) * 2 + 1];
        ++ctr;
        return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
    }
};

struct MtLib {                        // reference RNG, src/mtrng/random.cc behind bind(C)
    void* h = nullptr; void* rng = nullptr;
    void* (*fnew)() = nullptr; void (*fseed)(void*, unsigned) = nullptr;
    double (*frand)(void*) = nullptr; int (*frandint)(void*, int, int) = nullptr;
};

// ------------------------------------------------------------------ static tables
struct Tables {
    int nsst = 0, nb = 0, nops = 0, nstates = 0, dmax = 0;
    std::vector<int> dim, offset, connect;        // connect[sst*nops + (b*2+s)*2+(ca-1)]
    std::vector<long long> op_off;
    std::vector<double> opdata, evals;
    std::vector<int> first_state, state2sst;      // lowest Fock state of each block; Fock state -> block
    double egs = 0;
    int conn(int sst, int b, int s, int ca) const { return connect[(size_t)sst * nops + (b * 2 + s) * 2 + (ca - 1)]; }
    const double* op(int sst, int b, int s, int ca) const { return &opdata[op_off[(size_t)sst * nops + (b * 2 + s) * 2 + (ca - 1)]]; }
};

struct Op { double tau; int b, s, ca; };          // b,s 0-based; ca: 1 = creator, 2 = annihilator

struct Counters {
    double TryAdd = 0, TryRem = 0, TryGlob = 0, TryShift = 0;
    double AccAdd = 0, AccRem = 0, AccGlob = 0, AccShift = 0;
    double AccQNAdd = 0, AccQNRem = 0, AccQNGlob = 0;
    double TryAdd4 = 0, AccAdd4 = 0, TryRem4 = 0, AccRem4 = 0;
};

const int LOGW = 20;   // doubles per replay-log record

struct Chain {
    Tables T;
    // parameters
    double beta = 1, taudiff_max = 1, pglob = 0.005, pshift = 0.005, p4op = 0, pouter = 0;
    int nb = 1, nf = 2, nftau = 0, ntau = 1000, nlegmax = 100, maxhisto = 500, ncorr = 100, nslide = 1;
    bool has_szt = true; int meas_dm = 1;
    int nsym = 0; std::vector<int> symmoves;       // user symmetry moves and their inverses: SymMoves(2 NSymMove, 2 NBands), Trace.F90:403-404, 471-478
    std::vector<double> ftau;                      // (nb,2,nftau)  [ (b*2+s)*nftau + i ]
    // rng
    int rng_kind = 1; Philox ph; MtLib mt;
    // configuration
    std::vector<Op> ops;
    std::vector<std::vector<double>> M;            // per flavour f=b*2+s: k x k, column-major M(i,j)=M[i+j*k], i=creator idx
    std::vector<int> kf;
    int outer_sst = 0, window_position = 0, num_windows = 1;
    bool statesampling = false; int outer_state = 0;   // statesampling = 1: ONE outer eigenstate (0-based inside outer_sst) instead of the sum over the superstate
    double tauwin_min = 0, tauwin_max = 1;
    LogVal Trace{0, 1}, Det{0, 1};
    double cfgsign = -1;
    std::vector<LogVal> parttrace;
    Counters C;
    int error = 0;
    // accumulators
    std::vector<double> gtau, gleg, histo, dens, outer_hist, sign_hist, contrib_sst, outerstate_hist, signstate_hist, contrib_state;
    std::vector<long long> accpair;
    // offdiag = 1 (ORACLE ONLY so far): ONE matrix over all flavours, pair moves with independent flavours
    bool offdiag = false; std::vector<double> ftau_full;   // F(f_a, f_c, i) = Ftau_full(b_a, s_a, b_c, s_c, i) at [(f_a * nf + f_c) * nftau + i]
    std::vector<double> Mfull; int kfull = 0;             // inverse of the full hybridisation matrix, (creator idx, annihilator idx), flavour-major order
    std::vector<double> gtau_full;                        // [f_a][f_c][ntau]
    int ngiw = 0; std::vector<double> giw;         // MeasGiw: [nf][2 ngiw] complex sums (re, im), frequencies n = -ngiw .. ngiw - 1
    double meas_sign = 0; long long cnt_meas = 0, cnt_steps = 0;
    // replay log
    double* logp = nullptr; long long log_n = 0, log_cap = 0;
    double rec[LOGW];

    // ---------------------------------------------------------- rng
    double grnd() {
        if (rng_kind == 0) return mt.frand(mt.rng);
        return ph.next();
    }
    int randint(int lo, int hi) {
        if (rng_kind == 0) return mt.frandint(mt.rng, lo, hi);
        int r = lo + (int)(ph.next() * (double)(hi - lo + 1));
        return r > hi ? hi : r;
    }

    // ---------------------------------------------------------- F(tau): get_HybrF Trace.F90:1014-1033
    double hybrF(double tau, int b, int s) const {
        double sig = 1.0;
        if (tau < 0.0) { tau += beta; sig = -1.0; }
        double n = tau / beta * (double)(nftau - 1);
        int i = (int)n;
        const double* f = &ftau[(size_t)(b * 2 + s) * nftau];
        return sig * (f[i] + (n - (double)i) * (f[i + 1] - f[i]));
    }

    // ---------------------------------------------------------- window: Trace.F90:1608-1751
    void set_window_edges() {
        if (num_windows > 1) {
            tauwin_min = ((double)window_position / (double)(num_windows + 1)) * beta;
            tauwin_max = ((double)(window_position + 2) / (double)(num_windows + 1)) * beta;
        }
    }
    void init_window(double tdm) {
        ec.valid = false; fc.valid = false;
        window_position = 0;
        if (tdm > 0.0) {
            taudiff_max = tdm;
            num_windows = std::max((int)std::floor(beta / taudiff_max) - 1, 1);
            tauwin_min = ((double)window_position / (double)(num_windows + 1)) * beta;
            tauwin_max = ((double)(window_position + 2) / (double)(num_windows + 1)) * beta;
        } else {
            taudiff_max = beta; num_windows = 1; tauwin_min = 0; tauwin_max = beta;
        }
    }
    void shift_window() {                         // Trace.F90:1652-1678
        ec.valid = false;                            // (the per-operator caches do not depend on the window)
        if (num_windows > 1) {
            window_position += 1;
            if (window_position == num_windows) window_position = 0;
            set_window_edges();
        } else window_position = 0;
    }

    // ---------------------------------------------------------- NPairs: ensure_valid_NPairs_diag Trace.F90:1770-1815
    int count_pairs(const std::vector<Op>& o) const {
        if (offdiag) return count_pairs_offdiag(o);
        int np = 0;
        for (size_t i = 0; i < o.size(); ++i) {
            if (o[i].tau < tauwin_min) continue;
            if (o[i].tau > tauwin_max) break;
            for (size_t j = i + 1; j < o.size(); ++j) {
                if (o[j].b != o[i].b || o[j].s != o[i].s) continue;
                if (o[j].tau > tauwin_max) break;
                if (o[j].tau > o[i].tau + taudiff_max) break;
                if (o[j].ca != o[i].ca) ++np;
            }
        }
        return np;
    }

    // ---------------------------------------------------------- superstate sequence
    // sst after each operator starting from outer_sst (following_sst of Trace.F90:56); -1 = annihilated
    bool sst_sequence(const std::vector<Op>& o, int outer, std::vector<int>& after) const {
        after.resize(o.size());
        int cur = outer;
        for (size_t i = 0; i < o.size(); ++i) {
            cur = T.conn(cur, o[i].b, o[i].s, o[i].ca);
            if (cur < 0) return false;
            after[i] = cur;
        }
        return cur == outer;
    }

    // ---------------------------------------------------------- local trace: get_Trace_EB Trace.F90:5354-6033
    // (no caches: ket chain from tau=0 through all operators, closed at beta with <e_j|)
    static void split_log(double* v, int n, double& lf) {       // Trace.F90:5341-5348
        double mv = 0; for (int i = 0; i < n; ++i) mv = std::max(mv, std::fabs(v[i]));
        if (mv > 0) for (int i = 0; i < n; ++i) v[i] /= mv;
        lf += std::log(mv);
    }
    // only >= 0 (state sampling, louter_sst_size = 1 Trace.F90:5759-5764): the contribution of that single outer eigenstate
    LogVal trace_full(const std::vector<Op>& o, int outer, std::vector<LogVal>* parts, int only = -1) const {
        LogVal tot{-HUGE_D, 0.0};
        int m = T.dim[outer];
        if (parts) parts->assign(m, LogVal{-HUGE_D, 0.0});
        if (T.dmax == 1) {                                      // get_Trace_EB_onlyscalars 5556-5727
            int sst = outer; double lg = 0, sg = 1, tauR = 0;
            for (const Op& e : o) {
                int ns = T.conn(sst, e.b, e.s, e.ca);
                if (ns < 0) return tot;
                lg += -(T.evals[T.offset[sst]] - T.egs) * (e.tau - tauR);
                double el = T.op(sst, e.b, e.s, e.ca)[0];
                lg += std::log(std::fabs(el)); sg *= sgn(el);
                tauR = e.tau; sst = ns;
            }
            lg += -(T.evals[T.offset[sst]] - T.egs) * (beta - tauR);
            tot = LogVal{lg, sg};
            if (parts) (*parts)[0] = tot;
            return tot;
        }
        std::vector<double> kt(T.dmax), ktp(T.dmax);
        for (int j = (only >= 0 ? only : 0); j < (only >= 0 ? only + 1 : m); ++j) {                           // StateLoop 5765
            int sst = outer, ns_sst = m; double lg = 0, tauR = 0;
            std::fill(kt.begin(), kt.end(), 0.0); kt[j] = 1.0;
            bool dead = false;
            for (const Op& e : o) {
                int ns = T.conn(sst, e.b, e.s, e.ca);
                if (ns < 0) { dead = true; break; }
                int nn = T.dim[ns];
                const double* ev = &T.evals[T.offset[sst]];
                const double* A = T.op(sst, e.b, e.s, e.ca);     // (nn x ns_sst) column-major
                if (ns_sst > 1) {                               // 5802-5822
                    for (int i = 0; i < ns_sst; ++i) ktp[i] = std::exp(-(ev[i] - ev[0]) * (e.tau - tauR)) * kt[i];
                    lg += -(ev[0] - T.egs) * (e.tau - tauR);
                    split_log(ktp.data(), ns_sst, lg);
                    for (int r = 0; r < nn; ++r) kt[r] = 0.0;
                    for (int c = 0; c < ns_sst; ++c) { double x = ktp[c]; const double* col = A + (size_t)c * nn; for (int r = 0; r < nn; ++r) kt[r] += col[r] * x; }
                    split_log(kt.data(), nn, lg);
                } else {                                        // 5823-5829
                    lg += -(ev[0] - T.egs) * (e.tau - tauR);
                    double x = kt[0];
                    for (int r = 0; r < nn; ++r) kt[r] = x * A[r];
                }
                tauR = e.tau; sst = ns; ns_sst = nn;
            }
            if (dead) continue;
            // close with bra = e_j evolved from beta (6018-6031)
            const double* ev = &T.evals[T.offset[sst]];
            double bl = -(ev[0] - T.egs) * (beta - tauR);
            double braket = std::exp(-(ev[j] - ev[0]) * (beta - tauR)) * kt[j];   // split_log of a unit vector leaves it
            LogVal pt{lg + bl + std::log(std::fabs(braket)), sgn(braket)};
            if (braket == 0.0) pt = LogVal{-HUGE_D, 0.0};
            if (parts) (*parts)[j] = pt;
            if (pt.sign != 0.0) tot = (tot.sign == 0.0) ? pt : ladd(tot, pt);
        }
        return tot;
    }

    // ---------------------------------------------------------- windowed evaluation (opt-in; CPU baseline only)
    // The reference does not recompute the whole trace per proposal: operators outside the sliding window keep cached
    // states (TOper%state / %cache, update_trace_EB Trace.F90:6109-6209) and only the window is re-evaluated.  This mode
    // mirrors that cost structure for the timed CPU baseline: the ket chain of the operators left of the window and the bra
    // chain of the operators right of it are cached per outer state and rebuilt when the window moves or a global move /
    // tau shift is accepted; a pair move (always inside the window) multiplies through the window only.  Same products as
    // trace_full in a different association order; the checker keeps using trace_full (fast_trace = false).
    int fast_trace = 0;                              // 0 from scratch (checker), 1 cached window edges, 2 per-operator caches (reference cost structure)
    struct EdgeCache {
        bool valid = false; int wpos = -1, outer = -1, nL = 0, nR = 0, sstL = 0, sstR = 0; double tauL = 0, tauR = 0;
        std::vector<double> KL, lgL, BR, lgR;
    } ec;
    void build_edges(const std::vector<Op>& o, int outer) {
        const int n = (int)o.size(), m = T.dim[outer], dm = T.dmax;
        int nL = 0; while (nL < n && o[nL].tau < tauwin_min) ++nL;
        int kR = n; while (kR > nL && o[kR - 1].tau > tauwin_max) --kR;
        ec.nL = nL; ec.nR = n - kR; ec.wpos = window_position; ec.outer = outer; ec.valid = true;
        ec.KL.assign((size_t)m * dm, 0.0); ec.lgL.assign(m, 0.0); ec.BR.assign((size_t)m * dm, 0.0); ec.lgR.assign(m, 0.0);
        std::vector<double> kt(dm), ktp(dm);
        // kets of the operators left of the window (same arithmetic as trace_full)
        int sstL = outer; double tauL = 0;
        for (int j = 0; j < m; ++j) {
            int sst = outer, ns_sst = m; double lg = 0, tauR_ = 0;
            std::fill(kt.begin(), kt.end(), 0.0); kt[j] = 1.0;
            for (int i = 0; i < nL; ++i) {
                const Op& e = o[i];
                int ns = T.conn(sst, e.b, e.s, e.ca), nn = T.dim[ns];
                const double* ev = &T.evals[T.offset[sst]];
                const double* A = T.op(sst, e.b, e.s, e.ca);
                if (ns_sst > 1) {
                    for (int c = 0; c < ns_sst; ++c) ktp[c] = std::exp(-(ev[c] - ev[0]) * (e.tau - tauR_)) * kt[c];
                    lg += -(ev[0] - T.egs) * (e.tau - tauR_);
                    split_log(ktp.data(), ns_sst, lg);
                    for (int r = 0; r < nn; ++r) kt[r] = 0.0;
                    for (int c = 0; c < ns_sst; ++c) { double x = ktp[c]; const double* col = A + (size_t)c * nn; for (int r = 0; r < nn; ++r) kt[r] += col[r] * x; }
                    split_log(kt.data(), nn, lg);
                } else {
                    lg += -(ev[0] - T.egs) * (e.tau - tauR_);
                    double x = kt[0];
                    for (int r = 0; r < nn; ++r) kt[r] = x * A[r];
                }
                tauR_ = e.tau; sst = ns; ns_sst = nn;
            }
            for (int c = 0; c < ns_sst; ++c) ec.KL[(size_t)j * dm + c] = kt[c];
            ec.lgL[j] = lg; sstL = sst; tauL = tauR_;
        }
        ec.sstL = sstL; ec.tauL = tauL;
        // bras of the operators right of the window: b_j = e_j^T P O_n P ... O_kR, a vector in the block before operator kR
        std::vector<int> after; sst_sequence(o, outer, after);
        ec.sstR = kR == n ? outer : (kR == 0 ? outer : after[kR - 1]);
        ec.tauR = kR == n ? beta : o[kR].tau;
        for (int j = 0; j < m; ++j) {
            std::fill(kt.begin(), kt.end(), 0.0); kt[j] = 1.0;
            double lg = 0, tnext = beta; int sst = outer, ns_sst = m;          // block the bra lives in (after operator i)
            for (int i = n - 1; i >= kR; --i) {
                const Op& e = o[i];
                const int before = i == 0 ? outer : after[i - 1], nb_ = T.dim[before];
                const double* ev = &T.evals[T.offset[sst]];
                const double* A = T.op(before, e.b, e.s, e.ca);                 // (ns_sst x nb_) column-major
                for (int c = 0; c < ns_sst; ++c) ktp[c] = std::exp(-(ev[c] - ev[0]) * (tnext - e.tau)) * kt[c];
                lg += -(ev[0] - T.egs) * (tnext - e.tau);
                split_log(ktp.data(), ns_sst, lg);
                for (int c = 0; c < nb_; ++c) { const double* col = A + (size_t)c * ns_sst; double acc = 0; for (int r = 0; r < ns_sst; ++r) acc += col[r] * ktp[r]; kt[c] = acc; }
                split_log(kt.data(), nb_, lg);
                tnext = e.tau; sst = before; ns_sst = nb_;
            }
            for (int c = 0; c < ns_sst; ++c) ec.BR[(size_t)j * dm + c] = kt[c];
            ec.lgR[j] = lg;
        }
    }
    LogVal trace_windowed(const std::vector<Op>& o, int outer, std::vector<LogVal>* parts) {
        if (!ec.valid || ec.wpos != window_position || ec.outer != outer) build_edges(o, outer);
        const int n = (int)o.size(), m = T.dim[outer], dm = T.dmax;
        LogVal tot{-HUGE_D, 0.0};
        if (parts) parts->assign(m, LogVal{-HUGE_D, 0.0});
        std::vector<double> kt(dm), ktp(dm);
        const int lo = ec.nL, hi = n - ec.nR;
        for (int j = 0; j < m; ++j) {
            int sst = ec.sstL, ns_sst = T.dim[ec.sstL]; double lg = ec.lgL[j], tauR_ = ec.tauL;
            for (int c = 0; c < ns_sst; ++c) kt[c] = ec.KL[(size_t)j * dm + c];
            bool dead = false;
            for (int i = lo; i < hi; ++i) {
                const Op& e = o[i];
                int ns = T.conn(sst, e.b, e.s, e.ca);
                if (ns < 0) { dead = true; break; }
                int nn = T.dim[ns];
                const double* ev = &T.evals[T.offset[sst]];
                const double* A = T.op(sst, e.b, e.s, e.ca);
                if (ns_sst > 1) {
                    for (int c = 0; c < ns_sst; ++c) ktp[c] = std::exp(-(ev[c] - ev[0]) * (e.tau - tauR_)) * kt[c];
                    lg += -(ev[0] - T.egs) * (e.tau - tauR_);
                    split_log(ktp.data(), ns_sst, lg);
                    for (int r = 0; r < nn; ++r) kt[r] = 0.0;
                    for (int c = 0; c < ns_sst; ++c) { double x = ktp[c]; const double* col = A + (size_t)c * nn; for (int r = 0; r < nn; ++r) kt[r] += col[r] * x; }
                    split_log(kt.data(), nn, lg);
                } else {
                    lg += -(ev[0] - T.egs) * (e.tau - tauR_);
                    double x = kt[0];
                    for (int r = 0; r < nn; ++r) kt[r] = x * A[r];
                }
                tauR_ = e.tau; sst = ns; ns_sst = nn;
            }
            if (dead || sst != ec.sstR) continue;
            const double* ev = &T.evals[T.offset[sst]];
            double bl = -(ev[0] - T.egs) * (ec.tauR - tauR_), braket = 0.0;
            for (int c = 0; c < ns_sst; ++c) braket += ec.BR[(size_t)j * dm + c] * std::exp(-(ev[c] - ev[0]) * (ec.tauR - tauR_)) * kt[c];
            LogVal pt{lg + bl + ec.lgR[j] + std::log(std::fabs(braket)), sgn(braket)};
            if (braket == 0.0) pt = LogVal{-HUGE_D, 0.0};
            if (parts) (*parts)[j] = pt;
            if (pt.sign != 0.0) tot = (tot.sign == 0.0) ? pt : ladd(tot, pt);
        }
        return tot;
    }
    // ---------------------------------------------------------- per-operator state caches (opt-in, fast_trace = 2; CPU baseline)
    // The reference stores one ket or bra per operator (TOper%state / %cache, Trace.F90:47-49) and, for a proposal, restarts
    // from the innermost operators whose stored states are still valid (get_Trace_EB: StartR / StartL walk inward while
    // `.not. calc_new`, Trace.F90:5384-5470), calculates kets up to the changed operators and bras behind them, and swaps state
    // and cache on acceptance (update_trace_EB, Trace.F90:6109-6209).  Restated on the time-sorted array: KET_i (state after
    // operator i, one vector per outer state) is valid for i < kv, BRA_i (state before operator i, propagated down from beta)
    // for i >= bv; a proposal that changes the operators between old indices s and e extends the two frontiers to s-1 / e,
    // multiplies through the changed sub-string and closes with BRA_e.  Work per proposal ~ operators between the previous and
    // the current change, as in the reference; same products as trace_full in a different association order.
    struct Slot { std::vector<double> ket, bra, klg, blg; };
    struct FrontierCache { bool valid = false; int outer = -1, kv = 0, bv = 0; std::vector<Slot> slots; } fc;
    // v <- O_op exp(-(E-E0) dt) v for one outer state; returns false when the operator annihilates the block
    bool apply_ket(std::vector<double>& kt, std::vector<double>& ktp, int& sst, int& ns_sst, double& lg, const Op& e, double dt) const {
        int ns = T.conn(sst, e.b, e.s, e.ca);
        if (ns < 0) return false;
        int nn = T.dim[ns];
        const double* ev = &T.evals[T.offset[sst]];
        const double* A = T.op(sst, e.b, e.s, e.ca);
        if (ns_sst > 1) {
            for (int c = 0; c < ns_sst; ++c) ktp[c] = std::exp(-(ev[c] - ev[0]) * dt) * kt[c];
            lg += -(ev[0] - T.egs) * dt;
            split_log(ktp.data(), ns_sst, lg);
            for (int r = 0; r < nn; ++r) kt[r] = 0.0;
            for (int c = 0; c < ns_sst; ++c) { double x = ktp[c]; const double* col = A + (size_t)c * nn; for (int r = 0; r < nn; ++r) kt[r] += col[r] * x; }
            split_log(kt.data(), nn, lg);
        } else {
            lg += -(ev[0] - T.egs) * dt;
            double x = kt[0];
            for (int r = 0; r < nn; ++r) kt[r] = x * A[r];
        }
        sst = ns; ns_sst = nn;
        return true;
    }
    // b <- O_op^T exp(-(E-E0) dt) b : bra living after the operator (block `sst`) pulled in front of it (block `before`)
    void apply_bra(std::vector<double>& kt, std::vector<double>& ktp, int& sst, int& ns_sst, double& lg, const Op& e, int before, double dt) const {
        const int nb_ = T.dim[before];
        const double* ev = &T.evals[T.offset[sst]];
        const double* A = T.op(before, e.b, e.s, e.ca);                         // (ns_sst x nb_) column-major
        for (int c = 0; c < ns_sst; ++c) ktp[c] = std::exp(-(ev[c] - ev[0]) * dt) * kt[c];
        lg += -(ev[0] - T.egs) * dt;
        split_log(ktp.data(), ns_sst, lg);
        for (int c = 0; c < nb_; ++c) { const double* col = A + (size_t)c * ns_sst; double acc = 0; for (int r = 0; r < ns_sst; ++r) acc += col[r] * ktp[r]; kt[c] = acc; }
        split_log(kt.data(), nb_, lg);
        sst = before; ns_sst = nb_;
    }
    void fc_reset(int outer) {
        fc.valid = true; fc.outer = outer; fc.kv = 0; fc.bv = (int)ops.size();
        fc.slots.assign(ops.size(), Slot());
    }
    // local trace of the proposed configuration `nw` that differs from `ops` only between old indices s (first changed) and
    // e (first unchanged again); s == e for an insertion into one interval is allowed
    LogVal trace_cached(const std::vector<Op>& nw, int outer, int s, int e, std::vector<LogVal>* parts) {
        const int n = (int)ops.size(), m = T.dim[outer], dm = T.dmax, e_new = e + (int)nw.size() - n;
        if (!fc.valid || fc.outer != outer || (int)fc.slots.size() != n) fc_reset(outer);
        std::vector<int> after; sst_sequence(ops, outer, after);
        std::vector<double> kt(dm), ktp(dm);
        auto need = [&](std::vector<double>& v, size_t sz) { if (v.size() != sz) v.assign(sz, 0.0); };
        // kets of the unchanged head: KET_i for i in [kv, s)
        for (int i = fc.kv; i < s; ++i) {
            Slot& cur = fc.slots[i];
            need(cur.ket, (size_t)m * dm); need(cur.klg, m);
            const int sb = i == 0 ? outer : after[i - 1], db = T.dim[sb];
            const double tprev = i == 0 ? 0.0 : ops[i - 1].tau;
            for (int j = 0; j < m; ++j) {
                double lg; int sst = sb, ns_sst = db;
                if (i == 0) { std::fill(kt.begin(), kt.end(), 0.0); kt[j] = 1.0; lg = 0.0; }
                else { const Slot& pv = fc.slots[i - 1]; for (int c = 0; c < db; ++c) kt[c] = pv.ket[(size_t)j * dm + c]; lg = pv.klg[j]; }
                apply_ket(kt, ktp, sst, ns_sst, lg, ops[i], ops[i].tau - tprev);
                for (int c = 0; c < ns_sst; ++c) cur.ket[(size_t)j * dm + c] = kt[c];
                cur.klg[j] = lg;
            }
        }
        if (fc.kv < s) fc.kv = s;
        // bras of the unchanged tail: BRA_i for i in [e, bv), descending
        for (int i = fc.bv - 1; i >= e; --i) {
            Slot& cur = fc.slots[i];
            need(cur.bra, (size_t)m * dm); need(cur.blg, m);
            const int sa = after[i], da = T.dim[sa], before = i == 0 ? outer : after[i - 1];
            const double tnext = i + 1 == n ? beta : ops[i + 1].tau;
            for (int j = 0; j < m; ++j) {
                double lg; int sst = sa, ns_sst = da;
                if (i + 1 == n) { std::fill(kt.begin(), kt.end(), 0.0); kt[j] = 1.0; lg = 0.0; }
                else { const Slot& nx = fc.slots[i + 1]; for (int c = 0; c < da; ++c) kt[c] = nx.bra[(size_t)j * dm + c]; lg = nx.blg[j]; }
                apply_bra(kt, ktp, sst, ns_sst, lg, ops[i], before, tnext - ops[i].tau);
                for (int c = 0; c < ns_sst; ++c) cur.bra[(size_t)j * dm + c] = kt[c];
                cur.blg[j] = lg;
            }
        }
        if (fc.bv > e) fc.bv = e;
        // changed sub-string of the proposed configuration, closed with BRA_e
        LogVal tot{-HUGE_D, 0.0};
        if (parts) parts->assign(m, LogVal{-HUGE_D, 0.0});
        const int s_sst = s == 0 ? outer : after[s - 1];
        const double s_tau = s == 0 ? 0.0 : ops[s - 1].tau;
        const int e_sst = e == n ? outer : (e == 0 ? outer : after[e - 1]);
        const double e_tau = e == n ? beta : ops[e].tau;
        for (int j = 0; j < m; ++j) {
            int sst = s_sst, ns_sst = T.dim[s_sst]; double lg, tauR_ = s_tau;
            if (s == 0) { std::fill(kt.begin(), kt.end(), 0.0); kt[j] = 1.0; lg = 0.0; }
            else { const Slot& pv = fc.slots[s - 1]; for (int c = 0; c < ns_sst; ++c) kt[c] = pv.ket[(size_t)j * dm + c]; lg = pv.klg[j]; }
            bool dead = false;
            for (int i = s; i < e_new; ++i) {
                if (!apply_ket(kt, ktp, sst, ns_sst, lg, nw[i], nw[i].tau - tauR_)) { dead = true; break; }
                tauR_ = nw[i].tau;
            }
            if (dead || sst != e_sst) continue;
            const double* ev = &T.evals[T.offset[sst]];
            double bl = -(ev[0] - T.egs) * (e_tau - tauR_), braket = 0.0, blg = 0.0;
            if (e == n) braket = std::exp(-(ev[j] - ev[0]) * (e_tau - tauR_)) * kt[j];
            else { const Slot& nx = fc.slots[e]; blg = nx.blg[j];
                   for (int c = 0; c < ns_sst; ++c) braket += nx.bra[(size_t)j * dm + c] * std::exp(-(ev[c] - ev[0]) * (e_tau - tauR_)) * kt[c]; }
            LogVal pt{lg + bl + blg + std::log(std::fabs(braket)), sgn(braket)};
            if (braket == 0.0) pt = LogVal{-HUGE_D, 0.0};
            if (parts) (*parts)[j] = pt;
            if (pt.sign != 0.0) tot = (tot.sign == 0.0) ? pt : ladd(tot, pt);
        }
        return tot;
    }
    // an accepted pair move replaced old operators [s, e) by `nmid` new ones: hand the caches over (update_trace_EB)
    void fc_accept(int s, int e, int nmid) {
        if (!fc.valid) return;
        std::vector<Slot> ns; ns.reserve(fc.slots.size() + 2);
        for (int i = 0; i < s; ++i) ns.push_back(std::move(fc.slots[i]));
        for (int i = 0; i < nmid; ++i) ns.emplace_back();
        for (size_t i = e; i < fc.slots.size(); ++i) ns.push_back(std::move(fc.slots[i]));
        fc.slots.swap(ns);
        fc.kv = std::min(fc.kv, s); fc.bv = std::max(fc.bv, e) + (s + nmid - e);
    }
    // trace of a pair-move proposal: old operators [s, e) replaced (operators changed inside the window only)
    LogVal trace_pair(const std::vector<Op>& o, int outer, std::vector<LogVal>* parts, int s, int e) {
        if (statesampling) return trace_full(o, outer, parts, outer_state);
        if (fast_trace == 2 && T.dmax > 1) return trace_cached(o, outer, s, e, parts);
        if (fast_trace == 1 && T.dmax > 1 && num_windows > 1) return trace_windowed(o, outer, parts);
        return trace_full(o, outer, parts);
    }

    // ---------------------------------------------------------- signs: Trace.F90:3479-3618
    double get_sign(const std::vector<Op>& o, LogVal tr, LogVal det) const {
        std::vector<int> tn(nf, 0);
        for (const Op& e : o) tn[e.b * 2 + e.s]++;              // NOSOper
        std::vector<int> t1 = tn, t2 = tn;
        long long perm = 0, palt = 0;
        for (const Op& e : o) {
            // get_BlockSign 3594-3618
            t1[e.b * 2 + e.s]--;
            if (e.s == 1) { for (int b = 0; b <= e.b; ++b) perm += t1[b * 2 + 0]; }
            else          { for (int b = 0; b < e.b; ++b) perm += t1[b * 2 + 1]; }
            for (int b = 0; b < e.b; ++b) perm += t1[b * 2 + e.s];
            // get_permaltca 3554-3588
            t2[e.b * 2 + e.s]--;
            palt += t2[e.b * 2 + e.s] % (3 - e.ca);
        }
        if (((int)o.size() / 2) % 2 == 1) palt += 1;
        double bs = (perm % 2) ? -1.0 : 1.0, pa = (palt % 2) ? -1.0 : 1.0;
        return tr.sign * det.sign * bs * pa;
    }

    // ---------------------------------------------------------- dense LU helpers (DGETRF/DGETRI semantics, MatrixUpdate.F90:41-99, 302-339)
    static bool lu_decomp(std::vector<double>& a, int k, std::vector<int>& piv) {
        piv.resize(k);
        for (int c = 0; c < k; ++c) {
            int p = c; double mx = std::fabs(a[c + (size_t)c * k]);
            for (int r = c + 1; r < k; ++r) { double v = std::fabs(a[r + (size_t)c * k]); if (v > mx) { mx = v; p = r; } }
            piv[c] = p;
            if (mx == 0.0) return false;
            if (p != c) for (int j = 0; j < k; ++j) std::swap(a[c + (size_t)j * k], a[p + (size_t)j * k]);
            double d = a[c + (size_t)c * k];
            for (int r = c + 1; r < k; ++r) a[r + (size_t)c * k] /= d;
            for (int j = c + 1; j < k; ++j) { double x = a[c + (size_t)j * k]; if (x != 0.0) for (int r = c + 1; r < k; ++r) a[r + (size_t)j * k] -= a[r + (size_t)c * k] * x; }
        }
        return true;
    }
    static LogVal lu_logdet(const std::vector<double>& lu, int k, const std::vector<int>& piv) {
        LogVal d{0, 1};
        for (int i = 0; i < k; ++i) {
            double x = lu[i + (size_t)i * k];
            d.sign = (piv[i] != i) ? -d.sign * sgn(x) : d.sign * sgn(x);
            d.log += std::log(std::fabs(x));
        }
        return d;
    }
    // inverse + logdet of a (column-major k x k); singular -> log=-huge, sign=0
    static LogVal mat_logdet_inv(std::vector<double>& a, int k) {
        if (k == 0) return {0, 1};
        std::vector<double> lu = a; std::vector<int> piv;
        if (!lu_decomp(lu, k, piv)) return {-HUGE_D, 0.0};
        LogVal d = lu_logdet(lu, k, piv);
        // solve LU X = P I
        std::vector<double> inv((size_t)k * k, 0.0);
        for (int c = 0; c < k; ++c) {
            std::vector<double> b(k, 0.0); b[c] = 1.0;
            for (int i = 0; i < k; ++i) if (piv[i] != i) std::swap(b[i], b[piv[i]]);
            for (int i = 0; i < k; ++i) for (int j = 0; j < i; ++j) b[i] -= lu[i + (size_t)j * k] * b[j];
            for (int i = k - 1; i >= 0; --i) { for (int j = i + 1; j < k; ++j) b[i] -= lu[i + (size_t)j * k] * b[j]; b[i] /= lu[i + (size_t)i * k]; }
            for (int i = 0; i < k; ++i) inv[i + (size_t)c * k] = b[i];
        }
        a.swap(inv);
        return d;
    }
    static LogVal mat_logdet(const std::vector<double>& a, int k) {    // get_LogDetFull
        if (k == 1) return {std::log(std::fabs(a[0])), sgn(a[0])};
        std::vector<double> lu = a; std::vector<int> piv;
        if (!lu_decomp(lu, k, piv)) return {-HUGE_D, 0.0};
        return lu_logdet(lu, k, piv);
    }

    // ---------------------------------------------------------- off-diagonal hybridisation (offdiag = 1)
    double hybrF_full(double tau, int fa, int fc) const {        // get_HybrF_full Trace.F90:1038-1057
        double sig = 1.0;
        if (tau < 0.0) { tau += beta; sig = -1.0; }
        const double* F = &ftau_full[((size_t)fa * nf + fc) * nftau];
        double n = tau / beta * (double)(nftau - 1);
        int i = (int)n;
        if (i >= nftau - 1) return sig * F[nftau - 1];
        return sig * (F[i] + (n - (double)i) * (F[i + 1] - F[i]));
    }
    // operators in the order of get_FullHybr_offdiag (Trace.F90:2690-2770): by flavour (band outer, spin inner), then by time
    void flavour_major(const std::vector<Op>& o, std::vector<int>& cr, std::vector<int>& an) const {
        cr.clear(); an.clear();
        for (int f = 0; f < nf; ++f) for (size_t i = 0; i < o.size(); ++i) if (o[i].b * 2 + o[i].s == f) (o[i].ca == 1 ? cr : an).push_back((int)i);
    }
    // determinant of the full hybridisation matrix Hyb(annihilator, creator) and (optionally) its inverse (creator, annihilator)
    LogVal full_det(const std::vector<Op>& o, std::vector<double>* inv, int* kout) const {
        std::vector<int> cr, an; flavour_major(o, cr, an);
        if (cr.size() != an.size()) return {-HUGE_D, 0.0};
        int k = (int)cr.size();
        if (kout) *kout = k;
        if (k == 0) { if (inv) inv->clear(); return {0, 1}; }
        std::vector<double> a((size_t)k * k);
        for (int ic = 0; ic < k; ++ic) for (int ia = 0; ia < k; ++ia) {
            const Op& A = o[an[ia]]; const Op& Cc = o[cr[ic]];
            a[ia + (size_t)ic * k] = hybrF_full(A.tau - Cc.tau, A.b * 2 + A.s, Cc.b * 2 + Cc.s);
        }
        LogVal d = mat_logdet_inv(a, k);
        if (inv) *inv = a;
        return d;
    }
    // recalc_sign Trace.F90:3500-3546, literally: exchange "sort" of the keys with a swap count, plus sum_{i < n/2} i
    double recalc_sign(const std::vector<Op>& o) const {
        std::vector<double> key; key.reserve(o.size());
        for (const Op& e : o) key.push_back((double)(3 - e.ca) * 100000.0 + (double)(e.b + 1) * 10000.0 + (double)(e.s + 1) * 1000.0 + e.tau);
        long long perm = 0; const size_t n = key.size();
        for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) if (key[i] > key[j]) { std::swap(key[i], key[j]); ++perm; }
        for (long long i = 1; i <= (long long)n / 2 - 1; ++i) perm += i;
        return (perm % 2) ? -1.0 : 1.0;
    }
    double get_sign_offdiag(const std::vector<Op>& o, LogVal tr, LogVal det) const { return tr.sign * det.sign * recalc_sign(o); }   // Trace.F90:3483-3486
    int count_pairs_offdiag(const std::vector<Op>& o) const {       // ensure_valid_NPairs_offdiag / is_rempair_offdiag Trace.F90:2018-2030
        int np = 0;
        for (size_t i = 0; i < o.size(); ++i) {
            if (o[i].tau < tauwin_min) continue;
            if (o[i].tau > tauwin_max) break;
            for (size_t j = i + 1; j < o.size(); ++j) {
                if (o[j].tau > tauwin_max) break;
                if (o[j].tau > o[i].tau + taudiff_max) break;
                if (o[j].ca != o[i].ca) ++np;
            }
        }
        return np;
    }
    // StepAdd with b_offdiag (CTQMC.F90:3353-3612; pair_OperAdd force_diagonal = .false., Trace.F90:3657-3698): the creator draws its own
    // flavour; determinant from scratch (the b_full_offdiag branch, 3482-3487); (2 NBands)^2 / NPairs
    void StepAdd_offdiag() {
        log_begin(1);
        C.TryAdd += 1;
        double W = tauwin_max - tauwin_min;
        double tau_a = tauwin_min + W * grnd();
        int ba = randint(1, nb) - 1, sa = randint(1, 2) - 1;
        double tmin = std::max(tau_a - taudiff_max, tauwin_min);
        double w2 = std::min(tau_a + taudiff_max, tauwin_max) - tmin;
        double tau_c = tmin + w2 * grnd();
        int bc = randint(1, nb) - 1, sc = randint(1, 2) - 1;
        double taudiff_factor = W * w2;
        Op oa{tau_a, ba, sa, 2}, oc{tau_c, bc, sc, 1};
        rec[4] = (ba * 2 + sa) * nf + bc * 2 + sc;
        bool equal = (tau_a == tau_c);
        for (const Op& e : ops) if (e.tau == tau_a || e.tau == tau_c) equal = true;
        Op o1 = (tau_a < tau_c) ? oa : oc, o2 = (tau_a < tau_c) ? oc : oa;
        rec[2] = o1.tau; rec[3] = o2.tau;
        if (equal) { C.AccQNAdd += 1; log_end(2); return; }
        size_t i1 = 0; while (i1 < ops.size() && ops[i1].tau < o1.tau) ++i1;
        size_t i2 = i1; while (i2 < ops.size() && ops[i2].tau < o2.tau) ++i2;
        std::vector<Op> nw; nw.reserve(ops.size() + 2);
        nw.insert(nw.end(), ops.begin(), ops.begin() + i1); nw.push_back(o1);
        nw.insert(nw.end(), ops.begin() + i1, ops.begin() + i2); nw.push_back(o2);
        nw.insert(nw.end(), ops.begin() + i2, ops.end());
        {   // check_sst_sequence over the changed sub-string
            std::vector<int> after; sst_sequence(ops, outer_sst, after);
            int begin_sst = (i1 == 0) ? outer_sst : after[i1 - 1];
            int end_sst = (i2 == ops.size() || i2 == 0) ? outer_sst : after[i2 - 1];
            int cur = T.conn(begin_sst, o1.b, o1.s, o1.ca);
            for (size_t i = i1; i < i2 && cur >= 0; ++i) cur = T.conn(cur, ops[i].b, ops[i].s, ops[i].ca);
            if (cur >= 0) cur = T.conn(cur, o2.b, o2.s, o2.ca);
            if (cur < 0 || cur != end_sst) { C.AccQNAdd += 1; log_end(2); return; }
        }
        std::vector<double> Mn; int kn = 0;
        LogVal DetNew = full_det(nw, &Mn, &kn);
        LogVal DetRat = ldiv(DetNew, Det);
        {   // the same ratio by the rank-1 form on the stored inverse (get_LogDetRatPairAdd_full, Trace.F90:2835-2934): S = F_new - R M Q over ALL
            // flavours, sign (-1)^(row + column) of the new pair in the flavour-major order -- what an engine kernel would compute; logged
            // beside the from-scratch ratio (columns 18, 19) and compared by tests/test_oracle_offdiag.py
            std::vector<int> cr, an; flavour_major(ops, cr, an);
            const int k = (int)cr.size(), fa = ba * 2 + sa, fc = bc * 2 + sc;
            double S = hybrF_full(tau_a - tau_c, fa, fc);
            std::vector<double> Q(k), R(k);
            for (int j = 0; j < k; ++j) { const Op& A = ops[an[j]]; Q[j] = hybrF_full(A.tau - tau_c, A.b * 2 + A.s, fc); }
            for (int i = 0; i < k; ++i) { const Op& Cc = ops[cr[i]]; R[i] = hybrF_full(tau_a - Cc.tau, fa, Cc.b * 2 + Cc.s); }
            for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) S -= R[i] * Mfull[i + (size_t)j * k] * Q[j];
            int pc = 0, pa = 0;                                   // rank of the new creator / annihilator in the flavour-major order
            for (int i = 0; i < k; ++i) { const Op& Cc = ops[cr[i]]; const int f = Cc.b * 2 + Cc.s; if (f < fc || (f == fc && Cc.tau < tau_c)) ++pc; }
            for (int j = 0; j < k; ++j) { const Op& A = ops[an[j]]; const int f = A.b * 2 + A.s; if (f < fa || (f == fa && A.tau < tau_a)) ++pa; }
            rec[18] = std::log(std::fabs(S)); rec[19] = sgn(S) * (((pc + pa) % 2) ? -1.0 : 1.0);
        }
        std::vector<LogVal> parts;
        LogVal TraceNew = trace_full(nw, outer_sst, &parts);
        double rand = grnd();
        int NPairs = count_pairs_offdiag(nw);
        LogVal trr = ldiv(TraceNew, Trace);
        double prob = (DetNew.sign == 0.0) ? 0.0 : std::fabs(taudiff_factor * (double)(2 * nb) * (double)(2 * nb) / (double)NPairs * std::exp(trr.log + DetRat.log));
        rec[7] = DetRat.log; rec[8] = DetRat.sign; rec[9] = trr.log; rec[10] = trr.sign; rec[11] = prob; rec[14] = NPairs; rec[17] = rand;
        if (rand < prob) {
            ops.swap(nw); Mfull.swap(Mn); kfull = kn;
            Det = DetNew; Trace = TraceNew; parttrace = parts;
            cfgsign = -get_sign_offdiag(ops, Trace, Det);
            C.AccAdd += 1; log_end(1);
        } else log_end(0);
    }
    // StepRem with b_offdiag (CTQMC.F90:3892-4150; gen_OperRemove walks `next` instead of `fnext`, Trace.F90:4885-4905)
    void StepRem_offdiag() {
        log_begin(2);
        C.TryRem += 1;
        int NPairs = count_pairs_offdiag(ops);
        rec[14] = NPairs;
        if (NPairs <= 0) { C.AccQNRem += 1; log_end(2); return; }
        int NPair = randint(1, NPairs);
        size_t i1 = 0, i2 = 0; bool found = false;
        for (size_t i = 0; i < ops.size() && !found; ++i) {
            if (ops[i].tau < tauwin_min) continue;
            for (size_t j = i + 1; j < ops.size(); ++j) {
                if (ops[j].tau > tauwin_max) break;
                if (ops[j].tau > ops[i].tau + taudiff_max) break;
                if (ops[j].ca != ops[i].ca) { if (--NPair <= 0) { i1 = i; i2 = j; found = true; break; } }
            }
        }
        if (!found) { error = 2; log_end(2); return; }
        Op o1 = ops[i1], o2 = ops[i2];
        rec[2] = o1.tau; rec[3] = o2.tau;
        double tau_an = (o1.ca == 2) ? o1.tau : o2.tau;
        double factor = (double)NPairs / ((tauwin_max - tauwin_min)
                        * (std::min(tau_an + taudiff_max, tauwin_max) - std::max(tau_an - taudiff_max, tauwin_min)));
        std::vector<Op> nw; nw.reserve(ops.size());
        for (size_t i = 0; i < ops.size(); ++i) if (i != i1 && i != i2) nw.push_back(ops[i]);
        {
            std::vector<int> after; sst_sequence(ops, outer_sst, after);
            int begin_sst = (i1 == 0) ? outer_sst : after[i1 - 1];
            int end_sst = (i2 + 1 == ops.size()) ? outer_sst : after[i2];
            int cur = begin_sst;
            for (size_t i = i1 + 1; i < i2 && cur >= 0; ++i) cur = T.conn(cur, ops[i].b, ops[i].s, ops[i].ca);
            if (cur < 0 || cur != end_sst) { C.AccQNRem += 1; log_end(2); return; }
        }
        std::vector<double> Mn; int kn = 0;
        LogVal DetNew = full_det(nw, &Mn, &kn);
        LogVal DetRat = ldiv(DetNew, Det);
        {   // rank-1 form of the removal (get_LogDetRatPairRem_full, Trace.F90:2966-2997): the element of the stored inverse at (creator, annihilator)
            std::vector<int> cr, an; flavour_major(ops, cr, an);
            const int k = (int)cr.size();
            const size_t ic_op = o1.ca == 1 ? i1 : i2, ia_op = o1.ca == 2 ? i1 : i2;
            int pc = 0, pa = 0;
            for (int i = 0; i < k; ++i) { if (cr[i] == (int)ic_op) pc = i; if (an[i] == (int)ia_op) pa = i; }
            const double S = Mfull[pc + (size_t)pa * k];
            rec[18] = std::log(std::fabs(S)); rec[19] = sgn(S) * (((pc + pa) % 2) ? -1.0 : 1.0);
        }
        std::vector<LogVal> parts;
        LogVal TraceNew = trace_full(nw, outer_sst, &parts);
        LogVal trr = ldiv(TraceNew, Trace);
        double prob = (DetNew.sign == 0.0) ? 0.0 : std::fabs(factor / ((double)(2 * nb) * (double)(2 * nb)) * std::exp(trr.log + DetRat.log));
        double rand = grnd();
        rec[7] = DetRat.log; rec[8] = DetRat.sign; rec[9] = trr.log; rec[10] = trr.sign; rec[11] = prob; rec[17] = rand;
        if (rand < prob) {
            ops.swap(nw); Mfull.swap(Mn); kfull = kn;
            Det = DetNew; Trace = TraceNew; parttrace = parts;
            cfgsign = -get_sign_offdiag(ops, Trace, Det);
            C.AccRem += 1; log_end(1);
        } else log_end(0);
    }
    // MeasGtauRem_full CTQMC.F90:2475-2584: Gtau_full(f_a, f_c, bin) += +-M(c, a) cfgsign over ALL creator / annihilator pairs
    void measure_gtau_full() {
        std::vector<int> cr, an; flavour_major(ops, cr, an);
        const int k = (int)cr.size();
        for (int ic = 0; ic < k; ++ic) for (int ia = 0; ia < k; ++ia) {
            const Op& A = ops[an[ia]]; const Op& Cc = ops[cr[ic]];
            double Mv = Mfull[ic + (size_t)ia * k], dtau = A.tau - Cc.tau; int idt;
            if (dtau > 0.0) idt = (int)(dtau / beta * (double)ntau) + 1;
            else { idt = (int)((beta + dtau) / beta * (double)ntau) + 1; Mv = -Mv; }
            if (idt > ntau) idt = ntau;
            gtau_full[((size_t)(A.b * 2 + A.s) * nf + (Cc.b * 2 + Cc.s)) * ntau + idt - 1] += Mv * cfgsign;
        }
    }

    // Delta matrices from scratch: get_InvFullHybr Trace.F90:2279-2333 -> Mat(annih idx, creator idx)
    LogVal build_inverse_all(const std::vector<Op>& o, std::vector<std::vector<double>>& Mout, std::vector<int>& kout) const {
        Mout.assign(nf, {}); kout.assign(nf, 0);
        LogVal det{0, 1};
        for (int f = 0; f < nf; ++f) {
            std::vector<double> tc, ta;
            for (const Op& e : o) if (e.b * 2 + e.s == f) (e.ca == 1 ? tc : ta).push_back(e.tau);
            int k = (int)tc.size();
            kout[f] = k;
            if (k == 0 || ta.size() != tc.size()) { if (ta.size() != tc.size()) return {-HUGE_D, 0.0}; continue; }
            std::vector<double> a((size_t)k * k);
            for (int ic = 0; ic < k; ++ic) for (int ia = 0; ia < k; ++ia) a[ia + (size_t)ic * k] = hybrF(ta[ia] - tc[ic], f / 2, f % 2);
            LogVal d = mat_logdet_inv(a, k);       // inverse is indexed (creator idx, annih idx)
            det = lmul(det, d);
            Mout[f] = a;
        }
        return det;
    }

    // ---------------------------------------------------------- moves
    void log_begin(int type) { for (int i = 0; i < LOGW; ++i) rec[i] = 0; rec[0] = type; rec[4] = -1; }
    void log_end(int status) {
        rec[1] = status; rec[12] = (double)ops.size(); rec[13] = outer_sst; rec[15] = cfgsign;
    }

    // FPos helper: 1-based rank among same flavour/type with tau <= given (new op included by caller)
    int rank_of(const std::vector<Op>& o, int b, int s, int ca, double tau) const {
        int r = 0; for (const Op& e : o) if (e.b == b && e.s == s && e.ca == ca && e.tau < tau) ++r; return r + 1;
    }

    // WeightedOuterStateChoice Trace.F90:1477-1604 (state sampling; no outer truncation: every superstate keeps all its states).
    // Picks the new outer (superstate, eigenstate) among the superstates of `news` with probability ~ exp(-tau_new (E - Egs)) and
    // returns p_prop(new -> old) / p_prop(old -> new); entries < 0 are skipped; one random number, none when nothing can be chosen.
    double outer_tau(const std::vector<Op>& o) const { return o.empty() ? beta : beta - o.back().tau + o.front().tau; }
    double wosc(double tau_old, double tau_new, int sst_old, int state_old, const std::vector<int>& olds, const std::vector<int>& news,
                int& sst_new, int& state_new) {
        double tw = 0;
        for (int q : olds) if (q >= 0) for (int j = 0; j < T.dim[q]; ++j) tw += std::exp(-tau_old * (T.evals[T.offset[q] + j] - T.egs));
        double res = std::exp(-tau_old * (T.evals[T.offset[sst_old] + state_old] - T.egs)) / tw;
        sst_new = sst_old; state_new = state_old;
        double tot = 0; bool any = false;
        for (int q : news) if (q >= 0) for (int j = 0; j < T.dim[q]; ++j) { tot += std::exp(-tau_new * (T.evals[T.offset[q] + j] - T.egs)); any = true; }
        if (!any || tot <= 0.0) return 0.0;
        double rand = grnd() * tot, acc = 0;
        for (int q : news) if (q >= 0) for (int j = 0; j < T.dim[q]; ++j) {
            double w = std::exp(-tau_new * (T.evals[T.offset[q] + j] - T.egs));
            acc += w;
            if (rand <= acc && w != 0.0) { sst_new = q; state_new = j; return res * (tot / w); }
        }
        return 0.0;
    }

    // change_outer (PercentageOuterMove, CTQMC.F90:3422-3449 / 3934-3961): the superstates BETWEEN the two operators are kept and
    // the part of the sequence that runs from the later operator over beta to the earlier one is re-walked; the superstate it
    // passes tau = beta with becomes the outer superstate (check_sst_sequence Trace.F90:1278-1295), and the local trace is
    // evaluated from scratch because the outer superstate changed (CalcNew, Trace.F90:5381).
    void StepAdd(bool outer = false) {               // CTQMC.F90:3353-3612
        log_begin(outer ? 5 : 1);
        C.TryAdd += 1;
        // pair_OperAdd Trace.F90:3657-3698
        double W = tauwin_max - tauwin_min;
        double tau_a = tauwin_min + W * grnd();
        int b = randint(1, nb) - 1;
        int s = randint(1, 2) - 1;
        double tmin = std::max(tau_a - taudiff_max, tauwin_min);
        double w2 = std::min(tau_a + taudiff_max, tauwin_max) - tmin;
        double tau_c = tmin + w2 * grnd();
        double taudiff_factor = W * w2;
        Op oa{tau_a, b, s, 2}, oc{tau_c, b, s, 1};
        rec[4] = b * 2 + s;
        // process_OperAdd Trace.F90:4275-4462
        bool equal = (tau_a == tau_c);
        for (const Op& e : ops) if (e.tau == tau_a || e.tau == tau_c) equal = true;
        Op o1 = (tau_a < tau_c) ? oa : oc, o2 = (tau_a < tau_c) ? oc : oa;
        rec[2] = o1.tau; rec[3] = o2.tau;
        if (equal) { C.AccQNAdd += 1; log_end(2); return; }
        size_t i1 = 0; while (i1 < ops.size() && ops[i1].tau < o1.tau) ++i1;
        size_t i2 = i1; while (i2 < ops.size() && ops[i2].tau < o2.tau) ++i2;
        std::vector<Op> nw; nw.reserve(ops.size() + 2);
        nw.insert(nw.end(), ops.begin(), ops.begin() + i1); nw.push_back(o1);
        nw.insert(nw.end(), ops.begin() + i1, ops.begin() + i2); nw.push_back(o2);
        nw.insert(nw.end(), ops.begin() + i2, ops.end());
        int FPos[3];
        FPos[1] = rank_of(nw, b, s, 1, tau_c);
        FPos[2] = rank_of(nw, b, s, 2, tau_a);
        rec[5] = FPos[1]; rec[6] = FPos[2];
        // quantum-number check of the changed sub-string: check_sst_sequence Trace.F90:1243-1302
        int new_outer = outer_sst;
        if (!outer) {   // begin = Oper(1)%prev, end = Oper(2)%next (CTQMC.F90:3435-3438), ssts from the stored sequence
            std::vector<int> after; sst_sequence(ops, outer_sst, after);
            int begin_sst = (i1 == 0) ? outer_sst : after[i1 - 1];
            int end_sst = (i2 == ops.size() || i2 == 0) ? outer_sst : after[i2 - 1];
            int cur = T.conn(begin_sst, o1.b, o1.s, o1.ca);
            for (size_t i = i1; i < i2 && cur >= 0; ++i) cur = T.conn(cur, ops[i].b, ops[i].s, ops[i].ca);
            if (cur >= 0) cur = T.conn(cur, o2.b, o2.s, o2.ca);
            if (cur < 0 || cur != end_sst) { C.AccQNAdd += 1; log_end(2); return; }
        } else {        // begin = Oper(2)%prev, end = Oper(1)%next (CTQMC.F90:3422-3449; adjacent insertions: both are the "inner" sst)
            std::vector<int> after; sst_sequence(ops, outer_sst, after);
            int s_b = (i2 == 0) ? outer_sst : after[i2 - 1];          // superstate the later operator acts on (kept inner part ends here)
            int s_a = (i1 == 0) ? outer_sst : after[i1 - 1];          // superstate the earlier operator has to produce
            int cur = T.conn(s_b, o2.b, o2.s, o2.ca);
            for (size_t i = i2; i < ops.size() && cur >= 0; ++i) cur = T.conn(cur, ops[i].b, ops[i].s, ops[i].ca);
            new_outer = cur;                                          // passes tau = beta: the new outer superstate
            for (size_t i = 0; i < i1 && cur >= 0; ++i) cur = T.conn(cur, ops[i].b, ops[i].s, ops[i].ca);
            if (cur >= 0) cur = T.conn(cur, o1.b, o1.s, o1.ca);
            if (cur < 0 || cur != s_a) { C.AccQNAdd += 1; log_end(2); return; }
        }
        // state sampling: the new outer eigenstate is drawn inside the new outer superstate (CTQMC.F90:3471-3480)
        double osf = 1.0; int new_state = outer_state;
        if (statesampling && outer) osf = wosc(outer_tau(ops), outer_tau(nw), outer_sst, outer_state, {outer_sst}, {new_outer}, new_outer, new_state);
        // get_LogDetRatPairAdd Trace.F90:2775-2832 + get_DetRatAdd MatrixUpdate.F90:392-407
        int f = b * 2 + s, k = kf[f];
        std::vector<double> Q(k), R(k);
        { int iq = 0, ir = 0;
          for (const Op& e : ops) if (e.b == b && e.s == s) {
              if (e.ca == 2) Q[iq++] = hybrF(e.tau - tau_c, b, s);
              else           R[ir++] = hybrF(tau_a - e.tau, b, s);
          } }
        double S = hybrF(tau_a - tau_c, b, s);
        const std::vector<double>& Mf = M[f];
        for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) S -= R[i] * Mf[i + (size_t)j * k] * Q[j];
        double sig = (std::abs(FPos[1] - FPos[2]) % 2) ? -1.0 : 1.0;
        LogVal DetRat{std::log(std::fabs(S)), sgn(S) * sig};
        LogVal DetNew = lmul(DetRat, Det);
        std::vector<LogVal> parts;
        LogVal TraceNew = outer ? trace_full(nw, new_outer, &parts, statesampling ? new_state : -1) : trace_pair(nw, outer_sst, &parts, (int)i1, (int)i2);
        double rand = grnd();
        int NPairs = count_pairs(nw);
        LogVal trr = ldiv(TraceNew, Trace);
        double prob = std::fabs(osf * taudiff_factor * (double)(2 * nb) / (double)NPairs * std::exp(trr.log + DetRat.log));
        if (osf == 0.0) { trr = LogVal{0.0, 0.0}; prob = 0.0; }      // no outer state could be chosen: rejected whatever the trace
        rec[7] = DetRat.log; rec[8] = DetRat.sign; rec[9] = trr.log; rec[10] = trr.sign; rec[11] = prob; rec[14] = NPairs; rec[17] = rand;
        if (rand < prob) {
            // get_MatAdd MatrixUpdate.F90:410-442
            int p1 = FPos[1] - 1, p2 = FPos[2] - 1, kn = k + 1;
            std::vector<double> N((size_t)kn * kn);
            N[p1 + (size_t)p2 * kn] = 1.0 / S;
            for (int i = 0; i < k; ++i) { double acc = 0; for (int j = 0; j < k; ++j) acc += Mf[i + (size_t)j * k] * Q[j];
                int in = i < p1 ? i : i + 1; N[in + (size_t)p2 * kn] = -acc / S; }
            for (int j = 0; j < k; ++j) { double acc = 0; for (int i = 0; i < k; ++i) acc += R[i] * Mf[i + (size_t)j * k];
                int jn = j < p2 ? j : j + 1; N[p1 + (size_t)jn * kn] = -acc / S; }
            for (int i = 0; i < k; ++i) { int in = i < p1 ? i : i + 1;
                for (int j = 0; j < k; ++j) { int jn = j < p2 ? j : j + 1;
                    N[in + (size_t)jn * kn] = Mf[i + (size_t)j * k] + S * N[in + (size_t)p2 * kn] * N[p1 + (size_t)jn * kn]; } }
            M[f].swap(N); kf[f] = kn;
            ops.swap(nw);
            if (outer) { outer_sst = new_outer; outer_state = new_state; ec.valid = false; fc.valid = false; }
            else if (fast_trace == 2) fc_accept((int)i1, (int)i2, (int)(i2 - i1) + 2);
            Det = DetNew; Trace = TraceNew; parttrace = parts;
            cfgsign = -get_sign(ops, Trace, Det);
            C.AccAdd += 1;
            int bin = std::min(ntau, (int)std::floor((o2.tau - o1.tau) / beta * (double)ntau) + 1);
            accpair[(size_t)(b * 2 + s) * ntau + bin - 1] += 1;
            log_end(1);
        } else log_end(0);
    }

    void StepRem(bool outer = false) {               // CTQMC.F90:3892-4150
        log_begin(outer ? 6 : 2);
        C.TryRem += 1;
        // gen_OperRemove Trace.F90:4851-4950
        int NPairs = count_pairs(ops);
        rec[14] = NPairs;
        if (NPairs <= 0) { C.AccQNRem += 1; log_end(2); return; }
        int NPair = randint(1, NPairs);
        size_t i1 = 0, i2 = 0; bool found = false;
        for (size_t i = 0; i < ops.size() && !found; ++i) {
            if (ops[i].tau < tauwin_min) continue;
            for (size_t j = i + 1; j < ops.size(); ++j) {
                if (ops[j].b != ops[i].b || ops[j].s != ops[i].s) continue;
                if (ops[j].tau > tauwin_max) break;
                if (ops[j].tau > ops[i].tau + taudiff_max) break;
                if (ops[j].ca != ops[i].ca) { if (--NPair <= 0) { i1 = i; i2 = j; found = true; break; } }
            }
        }
        if (!found) { error = 2; log_end(2); return; }
        Op o1 = ops[i1], o2 = ops[i2];
        int b = o1.b, s = o1.s, f = b * 2 + s;
        rec[2] = o1.tau; rec[3] = o2.tau; rec[4] = f;
        int FPos[3];
        FPos[o1.ca] = rank_of(ops, b, s, o1.ca, o1.tau);
        FPos[o2.ca] = rank_of(ops, b, s, o2.ca, o2.tau);
        rec[5] = FPos[1]; rec[6] = FPos[2];
        double tau_an = (o1.ca == 2) ? o1.tau : o2.tau;
        double factor = (double)NPairs / ((tauwin_max - tauwin_min)
                        * (std::min(tau_an + taudiff_max, tauwin_max) - std::max(tau_an - taudiff_max, tauwin_min)));
        std::vector<Op> nw; nw.reserve(ops.size());
        for (size_t i = 0; i < ops.size(); ++i) if (i != i1 && i != i2) nw.push_back(ops[i]);
        int new_outer = outer_sst;
        if (!outer) {   // begin = Oper(1)%prev, end = Oper(2)%next (CTQMC.F90:3947-3950)
            std::vector<int> after; sst_sequence(ops, outer_sst, after);
            int begin_sst = (i1 == 0) ? outer_sst : after[i1 - 1];
            int end_sst = (i2 + 1 == ops.size()) ? outer_sst : after[i2];
            int cur = begin_sst;
            for (size_t i = i1 + 1; i < i2 && cur >= 0; ++i) cur = T.conn(cur, ops[i].b, ops[i].s, ops[i].ca);
            if (cur < 0 || cur != end_sst) { C.AccQNRem += 1; log_end(2); return; }
        } else {        // begin = Oper(2)%prev, end = Oper(1)%next (CTQMC.F90:3934-3961): the inner part keeps its superstates
            std::vector<int> after; sst_sequence(ops, outer_sst, after);
            int cur = after[i2 - 1];                                  // superstate before the later operator (i2 > i1 >= 0)
            for (size_t i = i2 + 1; i < ops.size() && cur >= 0; ++i) cur = T.conn(cur, ops[i].b, ops[i].s, ops[i].ca);
            new_outer = cur;
            for (size_t i = 0; i < i1 && cur >= 0; ++i) cur = T.conn(cur, ops[i].b, ops[i].s, ops[i].ca);
            if (cur < 0 || cur != after[i1]) { C.AccQNRem += 1; log_end(2); return; }
        }
        double osf = 1.0; int new_state = outer_state;                  // state sampling, CTQMC.F90:3979-3988
        if (statesampling && outer) osf = wosc(outer_tau(ops), outer_tau(nw), outer_sst, outer_state, {outer_sst}, {new_outer}, new_outer, new_state);
        // get_LogDetRatPairRem Trace.F90:2937-2963, get_DetRatRem MatrixUpdate.F90:445-452
        int k = kf[f];
        const std::vector<double>& Mf = M[f];
        int p1 = FPos[1] - 1, p2 = FPos[2] - 1;
        double S = Mf[p1 + (size_t)p2 * k];
        double sig = (std::abs(FPos[1] - FPos[2]) % 2) ? -1.0 : 1.0;
        LogVal DetRat{std::log(std::fabs(S)), sgn(S) * sig};
        LogVal DetNew = lmul(DetRat, Det);
        std::vector<LogVal> parts;
        LogVal TraceNew = outer ? trace_full(nw, new_outer, &parts, statesampling ? new_state : -1) : trace_pair(nw, outer_sst, &parts, (int)i1, (int)i2 + 1);
        LogVal trr = ldiv(TraceNew, Trace);
        double prob = std::fabs(osf * factor / (double)(2 * nb) * std::exp(trr.log + DetRat.log));
        if (osf == 0.0) { trr = LogVal{0.0, 0.0}; prob = 0.0; }
        double rand = grnd();
        rec[7] = DetRat.log; rec[8] = DetRat.sign; rec[9] = trr.log; rec[10] = trr.sign; rec[11] = prob; rec[17] = rand;
        if (rand < prob) {
            if (std::fabs(lval(DetRat)) < 1e-12) { error = 1; }       // CTQMC.F90:4062 `stop`
            // get_MatRem MatrixUpdate.F90:455-482
            int kn = k - 1;
            std::vector<double> N((size_t)kn * kn);
            for (int i = 0; i < kn; ++i) { int io = i < p1 ? i : i + 1;
                for (int j = 0; j < kn; ++j) { int jo = j < p2 ? j : j + 1;
                    N[i + (size_t)j * kn] = Mf[io + (size_t)jo * k] - Mf[io + (size_t)p2 * k] * Mf[p1 + (size_t)jo * k] / Mf[p1 + (size_t)p2 * k]; } }
            M[f].swap(N); kf[f] = kn;
            ops.swap(nw);
            if (outer) { outer_sst = new_outer; outer_state = new_state; ec.valid = false; fc.valid = false; }
            else if (fast_trace == 2) fc_accept((int)i1, (int)i2 + 1, (int)(i2 - i1) - 1);
            Det = DetNew; Trace = TraceNew; parttrace = parts;
            cfgsign = -get_sign(ops, Trace, Det);
            C.AccRem += 1;
            int bin = std::min(ntau, (int)std::floor((o2.tau - o1.tau) / beta * (double)ntau) + 1);
            accpair[(size_t)f * ntau + bin - 1] += 1;
            log_end(1);
        } else log_end(0);
    }

    // ---------------------------------------------------------- 4-operator moves (diagonal hybridisation): CTQMC.F90:3615-3889, 4153-4430
    // Two pairs are inserted / removed at once (Percentage4OperatorMove); the determinant of the proposed configuration is
    // rebuilt from scratch (get_InvFullHybr, CTQMC.F90:3816-3817) and the proposal ratio sums over the orders in which the
    // inverse move could take the four operators out (rempair_factor).
    bool is_rempair(const Op& a, const Op& b) const {            // is_rempair_diag Trace.F90:2005-2030
        return (offdiag || (a.b == b.b && a.s == b.s)) && a.ca != b.ca && a.tau >= tauwin_min && a.tau <= tauwin_max && b.tau >= tauwin_min && b.tau <= tauwin_max
               && std::fabs(a.tau - b.tau) <= taudiff_max;
    }
    static std::vector<Op> merged(const std::vector<Op>& base, const std::vector<Op>& add) {
        std::vector<Op> r = base; r.insert(r.end(), add.begin(), add.end());
        std::stable_sort(r.begin(), r.end(), [](const Op& x, const Op& y) { return x.tau < y.tau; });
        return r;
    }
    static std::vector<Op> without(const std::vector<Op>& base, const std::vector<Op>& drop) {
        std::vector<Op> r;
        for (const Op& e : base) { bool d = false; for (const Op& x : drop) if (x.tau == e.tau && x.b == e.b && x.s == e.s && x.ca == e.ca) d = true; if (!d) r.push_back(e); }
        return r;
    }
    // gen_OperRemove's pair #NPair in the reference's enumeration order (Trace.F90:4886-4914)
    bool pick_pair(const std::vector<Op>& o, int NPair, size_t& i1, size_t& i2) const {
        for (size_t i = 0; i < o.size(); ++i) {
            if (o[i].tau < tauwin_min) continue;
            for (size_t j = i + 1; j < o.size(); ++j) {
                if (!offdiag && (o[j].b != o[i].b || o[j].s != o[i].s)) continue;      // offdiag: `next` instead of `fnext` (Trace.F90:4885-4905)
                if (o[j].tau > tauwin_max) break;
                if (o[j].tau > o[i].tau + taudiff_max) break;
                if (o[j].ca != o[i].ca) { if (--NPair <= 0) { i1 = i; i2 = j; return true; } }
            }
        }
        return false;
    }
    double remfactor(int NPairs, const Op& a, const Op& b) const {   // hybpairremfactor Trace.F90:4939-4948 (around the annihilator)
        double tau_an = (a.ca == 2) ? a.tau : b.tau;
        return (double)NPairs / ((tauwin_max - tauwin_min) * (std::min(tau_an + taudiff_max, tauwin_max) - std::max(tau_an - taudiff_max, tauwin_min)));
    }
    void finish4(const std::vector<Op>& nw, double factor, int np_new, double& acc_counter) {
        std::vector<std::vector<double>> Mn; std::vector<int> kn;
        std::vector<double> Mfn; int kfn = 0;
        LogVal DetNew, DetRat;
        if (offdiag) {                               // get_FullHybr_offdiag + get_MatLogDetFull (CTQMC.F90:3809-3811, 4339-4341)
            DetNew = full_det(nw, &Mfn, &kfn);
            DetRat = DetNew.sign == 0.0 ? LogVal{0.0, 0.0} : ldiv(DetNew, Det);
        } else {
        DetNew = build_inverse_all(nw, Mn, kn);
        LogVal DetOld{0, 1};                         // get_FullLogDet: LU of the stored inverses (CTQMC.F90:3817, 4345)
        for (int f = 0; f < nf; ++f) if (kf[f] > 0) DetOld = lmul(mat_logdet(M[f], kf[f]), DetOld);
        DetRat = lmul(DetNew, DetOld);
        }
        std::vector<LogVal> parts;
        LogVal TraceNew = trace_full(nw, outer_sst, &parts, statesampling ? outer_state : -1);
        LogVal trr = ldiv(TraceNew, Trace);
        double rand = grnd();
        double prob = (offdiag && DetNew.sign == 0.0) ? 0.0 : std::fabs(factor * std::exp(trr.log + DetRat.log));
        rec[7] = DetRat.log; rec[8] = DetRat.sign; rec[9] = trr.log; rec[10] = trr.sign; rec[11] = prob; rec[14] = np_new; rec[17] = rand;
        if (rand < prob) {
            ops = nw; Det = DetNew; Trace = TraceNew; parttrace = parts; ec.valid = false; fc.valid = false;
            if (offdiag) { Mfull.swap(Mfn); kfull = kfn; cfgsign = -get_sign_offdiag(ops, Trace, Det); }
            else { M.swap(Mn); kf = kn; cfgsign = -get_sign(ops, Trace, Det); }
            acc_counter += 1; log_end(1);
        } else log_end(0);
    }
    void StepAdd4() {                                // CTQMC.F90:3615-3889 (b_offdiag = .false.)
        log_begin(7);
        C.TryAdd4 += 1;
        double W = tauwin_max - tauwin_min, fac = 1.0, tau_ref = 0; Op o[4];
        for (int i = 0; i < 4; ++i) {                // pair_OperAdd with N = 4, force_diagonal (Trace.F90:3657-3698)
            o[i].ca = (i % 2 == 0) ? 2 : 1;
            if (i % 2 == 0) {
                tau_ref = tauwin_min + W * grnd(); o[i].tau = tau_ref; fac *= W;
                o[i].b = randint(1, nb) - 1; o[i].s = randint(1, 2) - 1;
            } else {
                double tmin = std::max(tau_ref - taudiff_max, tauwin_min), w2 = std::min(tau_ref + taudiff_max, tauwin_max) - tmin;
                o[i].tau = tmin + w2 * grnd(); fac *= w2;
                if (offdiag) { o[i].b = randint(1, nb) - 1; o[i].s = randint(1, 2) - 1; }    // force_diagonal = .false. (CTQMC.F90:3641-3642)
                else { o[i].b = o[i - 1].b; o[i].s = o[i - 1].s; }
            }
        }
        const int old_N = count_pairs(ops);          // ensure_valid_NPairs 3651
        bool equal = false;
        for (int i = 0; i < 4; ++i) { for (const Op& e : ops) if (e.tau == o[i].tau) equal = true; for (int j = 0; j < i; ++j) if (o[j].tau == o[i].tau) equal = true; }
        Op O1 = o[0].tau < o[1].tau ? o[0] : o[1], O2 = o[0].tau < o[1].tau ? o[1] : o[0];
        Op O3 = o[2].tau < o[3].tau ? o[2] : o[3], O4 = o[2].tau < o[3].tau ? o[3] : o[2];
        std::vector<Op> all = merged(ops, {O1, O2, O3, O4});
        rec[2] = all.front().tau; rec[3] = all.back().tau;
        { double tmn = std::min(O1.tau, O3.tau), tmx = std::max(O2.tau, O4.tau); rec[2] = tmn; rec[3] = tmx; }
        rec[4] = (O1.b * 2 + O1.s) * nf + (O3.b * 2 + O3.s);
        if (equal) { log_end(2); return; }
        const int N12 = count_pairs(merged(ops, {O1, O2})), N34 = count_pairs(merged(ops, {O3, O4})), Nall = count_pairs(all);
        double rf;
        if (is_rempair(O1, O4) && is_rempair(O2, O3)) {
            const int N23 = count_pairs(merged(ops, {O2, O3})), N14 = count_pairs(merged(ops, {O1, O4}));
            rf = (1.0 / Nall) * (1.0 / N12 + 1.0 / N34 + 1.0 / N14 + 1.0 / N23); fac /= 4.0;
        } else if (is_rempair(O1, O3) && is_rempair(O2, O4)) {
            const int N23 = count_pairs(merged(ops, {O2, O4})), N14 = count_pairs(merged(ops, {O1, O3}));
            rf = (1.0 / Nall) * (1.0 / N12 + 1.0 / N34 + 1.0 / N14 + 1.0 / N23); fac /= 4.0;
        } else { rf = (1.0 / Nall) * (1.0 / N12 + 1.0 / N34); fac /= 2.0; }
        (void)old_N;
        std::vector<int> after;
        if (!sst_sequence(all, outer_sst, after)) { log_end(2); return; }   // check_sst_sequence over the changed sub-string (the rest is unchanged)
        const double nfl = (double)((2 * nb) * (2 * nb));                      // (2 NBands)^2, with offdiag (2 NBands)^4 (CTQMC.F90:3823-3826)
        finish4(all, fac * (offdiag ? nfl * nfl : nfl) * rf, Nall, C.AccAdd4);
    }
    void StepRem4() {                                // CTQMC.F90:4153-4430 (b_offdiag = .false.)
        log_begin(8);
        C.TryRem4 += 1;
        const int old_N = count_pairs(ops);
        rec[14] = old_N;
        if (old_N <= 0) { log_end(2); return; }
        size_t i1, i2;
        if (!pick_pair(ops, randint(1, old_N), i1, i2)) { error = 2; log_end(2); return; }
        const Op O1 = ops[i1], O2 = ops[i2];
        const double f1 = remfactor(old_N, O1, O2);
        const std::vector<Op> c34 = without(ops, {O1, O2});
        const int N34 = count_pairs(c34);
        if (N34 <= 0) { log_end(2); return; }
        size_t j1, j2;
        if (!pick_pair(c34, randint(1, N34), j1, j2)) { error = 2; log_end(2); return; }
        const Op O3 = c34[j1], O4 = c34[j2];
        const double f2 = remfactor(N34, O3, O4);
        double taudiff = (double)old_N * (double)N34 / f1 / f2;
        const int N12 = count_pairs(without(ops, {O3, O4}));
        double rf;
        if (is_rempair(O1, O4) && is_rempair(O2, O3)) {
            const int N23 = count_pairs(without(ops, {O1, O4})), N14 = count_pairs(without(ops, {O2, O3}));
            rf = (1.0 / old_N) * (1.0 / N12 + 1.0 / N34 + 1.0 / N14 + 1.0 / N23); taudiff /= 4.0;
        } else if (is_rempair(O1, O3) && is_rempair(O2, O4)) {
            const int N23 = count_pairs(without(ops, {O1, O3})), N14 = count_pairs(without(ops, {O2, O4}));
            rf = (1.0 / old_N) * (1.0 / N12 + 1.0 / N34 + 1.0 / N14 + 1.0 / N23); taudiff /= 4.0;
        } else { rf = (1.0 / old_N) * (1.0 / N12 + 1.0 / N34); taudiff /= 2.0; }
        const std::vector<Op> none = without(ops, {O1, O2, O3, O4});
        rec[2] = std::min(O1.tau, O3.tau); rec[3] = std::max(O2.tau, O4.tau);
        rec[4] = (O1.b * 2 + O1.s) * nf + (O3.b * 2 + O3.s);
        std::vector<int> after;
        if (!sst_sequence(none, outer_sst, after)) { log_end(2); return; }
        const double nfl = (double)((2 * nb) * (2 * nb));                      // CTQMC.F90:4353-4356
        finish4(none, 1.0 / (offdiag ? nfl * nfl : nfl) / taudiff / rf, count_pairs(none), C.AccRem4);
    }

    // count compatible outer superstates: check_qn_all_ssts Trace.F90:1402-1433
    int compatible_ssts(const std::vector<Op>& o, std::vector<int>& ssts) const {
        ssts.resize(T.nsst); int cnt = 0; std::vector<int> after;
        for (int k = 0; k < T.nsst; ++k) { if (sst_sequence(o, k, after)) { ssts[k] = k; ++cnt; } else ssts[k] = -1; }
        return cnt;
    }

    void StepGlob() {                                // CTQMC.F90:4433-4554
        log_begin(3);
        C.TryGlob += 1;
        if (ops.empty()) { log_end(2); return; }
        std::vector<int> gu(nf + 1, 0);              // gu[1..nf] target flavour (bit index+1), gu[0] = c/c^dag flip
        double r = grnd();
        int type;
        if (r < 0.3) {                               // gen_SpinFlipUpdate Trace.F90:1150-1166
            type = 3; for (int i = 1; i <= nb; ++i) gu[i] = i + nb; for (int i = nb + 1; i <= 2 * nb; ++i) gu[i] = i - nb;
        } else if (r < 0.6) {                        // gen_CAFlip 1186-1198
            type = 4; gu[0] = 1; for (int i = 1; i <= nf; ++i) gu[i] = i;
        } else if (r < 0.8 && nsym > 0) {            // gen_SymUpdate 1170-1182 (CTQMC.F90:4463-4465)
            type = 2;
            int isym = randint(1, 2 * nsym);
            for (int i = 1; i <= nf; ++i) gu[i] = symmoves[(size_t)(isym - 1) * nf + i - 1];
        } else {                                     // gen_GlobalUpdate_new 1103-1123
            type = 1; for (int i = 1; i <= nf; ++i) gu[i] = i;
            for (int j = 1; j <= nf; ++j) { int i = randint(j, nf); std::swap(gu[j], gu[i]); }
            gu[0] = randint(0, 1);
        }
        rec[4] = type;
        // globalUpdate Trace.F90:2126-2188: flavour index i-1 = b + nb*s
        std::vector<Op> nw = ops;
        for (Op& e : nw) { int i = e.b + nb * e.s + 1; int g = gu[i]; e.b = (g - 1) % nb; e.s = (g - 1) / nb; if (gu[0] != 0) e.ca = 3 - e.ca; }
        double factor = 0; int new_outer = outer_sst, new_state = outer_state;
        std::vector<int> after;
        if (type == 3 && has_szt) {                  // perform_spinflip_check_qn 1306-1339
            int st = T.first_state[outer_sst];
            int dn = st & ((1 << nb) - 1), up = (st >> nb) & ((1 << nb) - 1);
            int ns = up | (dn << nb);
            new_outer = T.state2sst[ns];
            factor = sst_sequence(nw, new_outer, after) ? 1.0 : 0.0;
            if (factor != 0.0 && statesampling) {    // Trace.F90:1325-1332
                const int cand = new_outer;
                factor = wosc(outer_tau(ops), outer_tau(ops), outer_sst, outer_state, {outer_sst}, {cand}, new_outer, new_state);
            }
        } else {                                     // perform_global_check_qn 1345-1394
            std::vector<int> so, sn;
            int pold = compatible_ssts(ops, so);
            if (pold <= 0) { error = 3; log_end(2); return; }
            int pnew = compatible_ssts(nw, sn);
            if (pnew > 0 && statesampling) factor = wosc(outer_tau(ops), outer_tau(ops), outer_sst, outer_state, so, sn, new_outer, new_state);   // Trace.F90:1372-1379
            else if (pnew > 0) {
                factor = (double)pnew / (double)pold;
                int kk = randint(1, pnew);
                for (int i = 0; i < T.nsst; ++i) if (sn[i] != -1) { if (--kk <= 0) { new_outer = sn[i]; break; } }
            }
        }
        if (factor == 0.0) { C.AccQNGlob += 1; log_end(2); return; }
        std::vector<std::vector<double>> Mn; std::vector<int> kn;
        std::vector<double> Mfn; int kfn = 0;
        LogVal DetNew, DetRat;
        if (offdiag) {                               // CTQMC.F90:4491-4494: the relabelled full matrix from scratch
            DetNew = full_det(nw, &Mfn, &kfn);
            if (DetNew.sign == 0.0) { (void)grnd(); log_end(0); return; }   // singular relabelled matrix: weight 0, rejected (the accept draw is still consumed)
            DetRat = ldiv(DetNew, Det);
        } else {
        DetNew = build_inverse_all(nw, Mn, kn);
        LogVal DetOld{0, 1};                         // get_FullLogDet Trace.F90:2665-2685 (LU of the stored inverses)
        for (int f = 0; f < nf; ++f) if (kf[f] > 0) DetOld = lmul(mat_logdet(M[f], kf[f]), DetOld);
        DetRat = lmul(DetNew, DetOld);
        }
        std::vector<LogVal> parts;
        LogVal TraceNew = trace_full(nw, new_outer, &parts, statesampling ? new_state : -1);
        LogVal trr = ldiv(TraceNew, Trace);
        double prob = std::fabs(std::exp(trr.log + DetRat.log) * factor);
        double rand = grnd();
        rec[7] = DetRat.log; rec[8] = DetRat.sign; rec[9] = trr.log; rec[10] = trr.sign; rec[11] = prob; rec[17] = rand;
        if (rand < prob) {
            ops.swap(nw); outer_sst = new_outer; outer_state = new_state;
            if (offdiag) { Mfull.swap(Mfn); kfull = kfn; } else { M.swap(Mn); kf = kn; }
            Det = DetNew; Trace = TraceNew; parttrace = parts; ec.valid = false; fc.valid = false;
            cfgsign = offdiag ? -get_sign_offdiag(ops, Trace, Det) : -get_sign(ops, Trace, Det);
            C.AccGlob += 1; log_end(1);
        } else log_end(0);
    }

    // empty-trace choice of the outer (superstate, eigenstate) with state sampling: CTQMC.F90:4580-4608 / Trace.F90:612-643
    void choose_outer_state(double rand) {
        std::vector<double> w; double sum = 0;
        for (int i = 0; i < T.nsst; ++i) for (int j = 0; j < T.dim[i]; ++j) { w.push_back(lval(trace_full(ops, i, nullptr, j))); sum += w.back(); }
        double acc = 0; size_t q = 0;
        for (int i = 0; i < T.nsst; ++i) for (int j = 0; j < T.dim[i]; ++j, ++q) {
            outer_sst = i; outer_state = j; acc += w[q] / sum;
            if (acc >= rand && w[q] != 0.0) return;
        }
    }
    void StepChangeOuter() {                         // CTQMC.F90:4570-4638
        if (statesampling) {
            double rand = grnd();
            choose_outer_state(rand);
            Trace = trace_full(ops, outer_sst, &parttrace, outer_state); ec.valid = false; fc.valid = false;
            cfgsign = -get_sign(ops, Trace, Det);
            rec[17] = rand;
            return;
        }
        std::vector<double> w(T.nsst);
        double sum = 0;
        for (int i = 0; i < T.nsst; ++i) { w[i] = lval(trace_full(ops, i, nullptr)); sum += w[i]; }
        for (double& x : w) x /= sum;
        double rand = grnd(), acc = 0; int pick = T.nsst - 1;
        for (int i = 0; i < T.nsst; ++i) { acc += w[i]; pick = i; if (acc >= rand && w[i] != 0.0) break; }
        outer_sst = pick;
        Trace = trace_full(ops, outer_sst, &parttrace); ec.valid = false; fc.valid = false;
        cfgsign = -get_sign(ops, Trace, Det);
        rec[17] = rand;
    }

    void StepShiftTau() {                            // CTQMC.F90:4557-4835
        log_begin(4);
        if (ops.empty()) { StepChangeOuter(); log_end(1); return; }
        double dt = grnd() * beta;
        rec[2] = dt;
        if (dt == 0.0 || dt == beta) { log_end(2); return; }
        C.TryShift += 1;
        size_t n = ops.size();
        size_t nwrap = 0; while (nwrap < n && ops[nwrap].tau < dt) ++nwrap;   // operators passing through 0
        std::vector<int> wrapped(nf * 2, 0);
        for (size_t i = 0; i < nwrap; ++i) wrapped[(ops[i].b * 2 + ops[i].s) * 2 + (ops[i].ca - 1)]++;
        std::vector<int> after; sst_sequence(ops, outer_sst, after);
        // new first operator = first with tau >= dt (or the old first if all wrap); new outer = its preceding sst
        size_t first = (nwrap == n) ? 0 : nwrap;
        int new_outer = (first == 0) ? outer_sst : after[first - 1];
        std::vector<Op> nw; nw.reserve(n);
        for (size_t i = 0; i < n; ++i) { Op e = ops[(first + i) % n]; e.tau = std::fmod(e.tau - dt, beta); if (e.tau < 0) e.tau += beta; nw.push_back(e); }
        double osf = 1.0; int new_state = outer_state;                  // CTQMC.F90:4702-4710
        if (statesampling) { const int cand = new_outer; osf = wosc(outer_tau(ops), outer_tau(nw), outer_sst, outer_state, {outer_sst}, {cand}, new_outer, new_state); }
        int wp_old = window_position; double wmin_old = tauwin_min, wmax_old = tauwin_max;
        int sh = (int)std::floor((dt / beta) * (double)(num_windows + 1)) % num_windows;
        window_position = ((wp_old - sh) % num_windows + num_windows) % num_windows;
        set_window_edges();
        std::vector<LogVal> parts;
        LogVal TraceNew = osf > 0.0 ? trace_full(nw, new_outer, &parts, statesampling ? new_state : -1) : LogVal{-HUGE_D, 0.0};
        LogVal trr = osf > 0.0 ? ldiv(TraceNew, Trace) : LogVal{0.0, 0.0};
        double prob = osf > 0.0 ? std::fabs(osf * lval(trr)) : 0.0;
        double rand = grnd();
        rec[9] = trr.log; rec[10] = trr.sign; rec[11] = prob; rec[17] = rand;
        if (rand < prob && offdiag) {
            // offdiag: the rolled matrix (update_InvFullHybrShiftTau, Trace.F90:2336-2352 / 2460-2560) is rebuilt from scratch here -- same
            // matrix, same determinant (its modulus cannot change, its sign can)
            ops.swap(nw); outer_sst = new_outer; outer_state = new_state;
            Trace = TraceNew; parttrace = parts; Det = full_det(ops, &Mfull, &kfull); ec.valid = false; fc.valid = false;
            cfgsign = -get_sign_offdiag(ops, Trace, Det);
            C.AccShift += 1; log_end(1);
        } else
        if (rand < prob) {
            // update_InvFullHybrShiftTau_diag Trace.F90:2355-2457
            double detrat = 1.0;
            for (int f = 0; f < nf; ++f) {
                int k = kf[f], w1 = wrapped[f * 2 + 0], w2 = wrapped[f * 2 + 1];
                if (!((w1 == 0 && w2 == 0) || (w1 == k && w2 == k))) {
                    std::vector<double> N((size_t)k * k); const std::vector<double>& O = M[f];
                    for (int j = 0; j < k - w2; ++j) {
                        for (int i = 0; i < k - w1; ++i) N[i + (size_t)j * k] = O[(i + w1) + (size_t)(j + w2) * k];
                        for (int i = 0; i < w1; ++i) N[(i + k - w1) + (size_t)j * k] = -O[i + (size_t)(j + w2) * k];
                    }
                    for (int j = 0; j < w2; ++j) {
                        for (int i = 0; i < k - w1; ++i) N[i + (size_t)(j + k - w2) * k] = -O[(i + w1) + (size_t)j * k];
                        for (int i = 0; i < w1; ++i) N[(i + k - w1) + (size_t)(j + k - w2) * k] = O[i + (size_t)j * k];
                    }
                    M[f].swap(N);
                }
                if ((k % 2) != 0 && ((w1 + w2) % 2) != 0) detrat = -detrat;
            }
            ops.swap(nw); outer_sst = new_outer; outer_state = new_state;
            Trace = TraceNew; parttrace = parts; Det.sign *= detrat; ec.valid = false; fc.valid = false;
            cfgsign = -get_sign(ops, Trace, Det);
            C.AccShift += 1; log_end(1);
        } else {
            window_position = wp_old; tauwin_min = wmin_old; tauwin_max = wmax_old;
            log_end(0);
        }
    }

    void step() {                                    // CTQMC.F90:8213-8258 / 8428-8533
        double r = grnd();
        if (offdiag) {                               // ORACLE-ONLY groundwork for offdiag = 1: pair moves, 4-operator moves, global moves, tau shifts
            if (r < (1.0 - pshift - pglob) / 2.0) { if (grnd() >= p4op) { (void)grnd(); StepAdd_offdiag(); } else StepAdd4(); }     // draws of CTQMC.F90:8441-8442
            else if (r <= 1.0 - pshift - pglob) { if (grnd() >= p4op) { (void)grnd(); StepRem_offdiag(); } else StepRem4(); }
            else if (r <= 1.0 - pshift) StepGlob();
            else StepShiftTau();
        } else
        if (r < (1.0 - pshift - pglob) / 2.0) {
            if (grnd() >= p4op) StepAdd(grnd() < pouter); else StepAdd4();
        } else if (r <= 1.0 - pshift - pglob) {
            if (grnd() >= p4op) StepRem(grnd() < pouter); else StepRem4();
        } else if (r <= 1.0 - pshift) StepGlob();
        else StepShiftTau();
        long long att = (long long)(C.TryAdd + C.TryRem);
        if (att % nslide == 0) shift_window();
        rec[16] = window_position;
        if (logp && log_n < log_cap) { std::memcpy(logp + log_n * LOGW, rec, sizeof(rec)); ++log_n; }
        ++cnt_steps;
    }

    // ---------------------------------------------------------- measurements (CTQMC.F90:8541-8581)
    void measure() {
        ++cnt_meas;
        meas_sign += cfgsign;                                                  // MeasSign 3343
        // MeasOuterHistograms 3285-3339
        outer_hist[outer_sst] += 1.0; sign_hist[outer_sst] += cfgsign;
        if (statesampling) { int q = T.offset[outer_sst] + outer_state; double e = std::exp(parttrace[outer_state].log);
          outerstate_hist[q] += 1.0; signstate_hist[q] += cfgsign; contrib_state[q] += e; contrib_sst[outer_sst] += e; }
        else
        { int off = T.offset[outer_sst], m = T.dim[outer_sst]; double c = 0;
          for (int i = 0; i < m; ++i) { outerstate_hist[off + i] += 1.0;
              signstate_hist[off + i] += Trace.sign * parttrace[i].sign * cfgsign;
              double e = std::exp(parttrace[i].log); c += e; contrib_state[off + i] += e; }
          contrib_sst[outer_sst] += c; }
        if (offdiag) {                                                         // NOSOper / 2 per flavour, MeasGtauRem_full instead of MeasGtauRem
            std::vector<int> cnt(nf, 0); for (const Op& e : ops) cnt[e.b * 2 + e.s]++;
            for (int f = 0; f < nf; ++f) histo[(size_t)f * (maxhisto + 1) + std::min(cnt[f] / 2, maxhisto)] += 1.0;
            if (meas_dm) measure_density();
            measure_gtau_full();
            return;
        }
        for (int f = 0; f < nf; ++f) histo[(size_t)f * (maxhisto + 1) + std::min(kf[f], maxhisto)] += 1.0;   // MeasExpOrder 3264
        if (meas_dm) measure_density();
        // MeasGtauRem 2252-2440 (isDetRat = .false.)
        for (int f = 0; f < nf; ++f) {
            int k = kf[f]; if (k == 0) continue;
            std::vector<double> tc, ta;
            for (const Op& e : ops) if (e.b * 2 + e.s == f) (e.ca == 1 ? tc : ta).push_back(e.tau);
            // Matsubara estimator, lookup branch of MeasGtauRem (CTQMC.F90:2400-2414; the non-lookup measure_giw_nfft 1015-1056 sums the same
            // terms): Giw(n) += -cfgsign M(c, a) exp(i w_n (tau_a - tau_c)), w_n = (2n + 1) pi / beta -- summed directly instead of by NFFT
            if (ngiw > 0) for (int ia = 0; ia < k; ++ia) for (int ic = 0; ic < k; ++ic) {
                const double mv = -cfgsign * M[f][ic + (size_t)ia * k], x = (ta[ia] - tc[ic]) / beta;
                for (int m = 0; m < 2 * ngiw; ++m) {
                    const double ph = M_PI * (double)(2 * (m - ngiw) + 1) * x;
                    giw[2 * ((size_t)f * 2 * ngiw + m)] += mv * std::cos(ph); giw[2 * ((size_t)f * 2 * ngiw + m) + 1] += mv * std::sin(ph);
                }
            }
            for (int ia = 0; ia < k; ++ia) for (int ic = 0; ic < k; ++ic) {
                double Mv = M[f][ic + (size_t)ia * k];
                double dtau = ta[ia] - tc[ic], x; int idt;
                if (dtau > 0.0) { idt = (int)(dtau / beta * (double)ntau) + 1; x = 2.0 / beta * dtau - 1.0; }
                else { idt = (int)((beta + dtau) / beta * (double)ntau) + 1; x = 2.0 / beta * (beta + dtau) - 1.0; Mv = -Mv; }
                Mv *= cfgsign;
                double Plx = 0, Plm1 = 0, Pl = 0;
                for (int l = 0; l < nlegmax; ++l) {
                    if (l == 0) Pl = 1.0; else if (l == 1) Pl = x;
                    else { int ll = l - 1; Pl = (2.0 * ll + 1.0) / (double)(ll + 1) * x * Plx - (double)ll / (double)(ll + 1) * Plm1; }
                    gleg[(size_t)f * nlegmax + l] += Mv * Pl;
                    Plm1 = Plx; Plx = Pl;
                }
                if (idt > ntau) idt = ntau;
                gtau[(size_t)f * ntau + idt - 1] += Mv;
            }
        }
    }
    // update_b2_states_eb Trace.F90:6213-6469 + MeasDensityMatrix_beta_half CTQMC.F90:3088-3151
    void measure_density() {
        int ns = T.nstates;
        int m = T.dim[outer_sst];
        if (ops.empty()) {
            int off = T.offset[outer_sst];
            for (int j = (statesampling ? outer_state : 0); j < (statesampling ? outer_state + 1 : m); ++j) dens[(size_t)(off + j) * ns + off + j] += -cfgsign * lval(ldiv(parttrace[j], Trace));
            return;
        }
        double bh = beta / 2.0;
        std::vector<double> kt(T.dmax), bt(T.dmax), tmp(T.dmax);
        for (int j = (statesampling ? outer_state : 0); j < (statesampling ? outer_state + 1 : m); ++j) {
            // ket: apply operators with tau <= beta/2
            int sst = outer_sst, nst = m; double kl = 0, tauR = 0;
            std::fill(kt.begin(), kt.end(), 0.0); kt[j] = 1.0;
            size_t i = 0;
            for (; i < ops.size() && ops[i].tau <= bh; ++i) {
                const Op& e = ops[i]; int nsst = T.conn(sst, e.b, e.s, e.ca), nn = T.dim[nsst];
                const double* ev = &T.evals[T.offset[sst]]; const double* A = T.op(sst, e.b, e.s, e.ca);
                for (int q = 0; q < nst; ++q) tmp[q] = std::exp(-(ev[q] - ev[0]) * (e.tau - tauR)) * kt[q];
                kl += -(ev[0] - T.egs) * (e.tau - tauR); split_log(tmp.data(), nst, kl);
                for (int r = 0; r < nn; ++r) kt[r] = 0; for (int c = 0; c < nst; ++c) { double x = tmp[c]; for (int r = 0; r < nn; ++r) kt[r] += A[r + (size_t)c * nn] * x; }
                split_log(kt.data(), nn, kl);
                tauR = e.tau; sst = nsst; nst = nn;
            }
            { const double* ev = &T.evals[T.offset[sst]];
              for (int q = 0; q < nst; ++q) kt[q] = std::exp(-(ev[q] - ev[0]) * (bh - tauR)) * kt[q];
              split_log(kt.data(), nst, kl); kl += -(ev[0] - T.egs) * (bh - tauR); }
            int sst_k = sst, nk = nst;
            // bra: apply operators with tau > beta/2 from the top
            sst = outer_sst; nst = m; double bl = 0, tauL = beta;
            std::fill(bt.begin(), bt.end(), 0.0); bt[j] = 1.0;
            for (size_t p = ops.size(); p > i; --p) {
                const Op& e = ops[p - 1]; int nsst = T.conn(sst, e.b, e.s, 3 - e.ca), nn = T.dim[nsst];
                const double* ev = &T.evals[T.offset[sst]]; const double* A = T.op(sst, e.b, e.s, 3 - e.ca);
                for (int q = 0; q < nst; ++q) tmp[q] = std::exp(-(ev[q] - ev[0]) * (tauL - e.tau)) * bt[q];
                bl += -(ev[0] - T.egs) * (tauL - e.tau); split_log(tmp.data(), nst, bl);
                for (int r = 0; r < nn; ++r) bt[r] = 0; for (int c = 0; c < nst; ++c) { double x = tmp[c]; for (int r = 0; r < nn; ++r) bt[r] += A[r + (size_t)c * nn] * x; }
                split_log(bt.data(), nn, bl);
                tauL = e.tau; sst = nsst; nst = nn;
            }
            { const double* ev = &T.evals[T.offset[sst]];
              for (int q = 0; q < nst; ++q) bt[q] = std::exp((ev[q] - ev[0]) * (bh - tauL)) * bt[q];
              split_log(bt.data(), nst, bl); bl += (ev[0] - T.egs) * (bh - tauL); }
            if (sst != sst_k) { error = 4; return; }
            int off = T.offset[sst_k];
            double den = lval(ldiv(Trace, LogVal{kl + bl, -cfgsign}));
            for (int a = 0; a < nk; ++a) { double fac = bt[a] / den; for (int c = 0; c < nk; ++c) dens[(size_t)(off + a) * ns + off + c] += fac * kt[c]; }
        }
    }

    // ---------------------------------------------------------- init_Trace Trace.F90:340-701 (empty config, outer sst ~ empty-trace weight)
    void init_trace() {
        ops.clear(); M.assign(nf, {}); kf.assign(nf, 0);
        Det = {0, 1};
        if (statesampling) {
            choose_outer_state(grnd());
            Trace = trace_full(ops, outer_sst, &parttrace, outer_state); ec.valid = false; fc.valid = false;
            cfgsign = -1.0;
            return;
        }
        outer_state = 0;
        std::vector<double> w(T.nsst); double sum = 0;
        for (int i = 0; i < T.nsst; ++i) { w[i] = lval(trace_full(ops, i, nullptr)); sum += w[i]; }
        for (double& x : w) x /= sum;
        double rand = grnd(), acc = 0; int pick = T.nsst - 1;
        for (int i = 0; i < T.nsst; ++i) { acc += w[i]; pick = i; if (acc >= rand && w[i] != 0.0) break; }
        outer_sst = pick;
        Trace = trace_full(ops, outer_sst, &parttrace); ec.valid = false; fc.valid = false;
        cfgsign = -1.0;                              // TTrace default (Trace.F90:129), first computed on the first accepted move
    }
    void set_config(int n, const double* taus, const int* orbs, const int* spins, const int* cas, int osst, int ostate = 0) {
        // set_mc_config_from_arrays Trace.F90:6872-7000
        ops.clear();
        for (int i = 0; i < n; ++i) ops.push_back(Op{taus[i], orbs[i] - 1, spins[i] - 1, cas[i]});
        outer_sst = osst; outer_state = statesampling ? ostate : 0; window_position = 0; set_window_edges();
        Trace = trace_full(ops, outer_sst, &parttrace, statesampling ? outer_state : -1); ec.valid = false; fc.valid = false;
        Det = build_inverse_all(ops, M, kf);
        cfgsign = -get_sign(ops, Trace, Det);
    }
    void reset_counters() { C = Counters(); std::fill(accpair.begin(), accpair.end(), 0); }   // init_counters CTQMC.F90:847-899
    void alloc_accumulators() {
        gtau.assign((size_t)nf * ntau, 0); gleg.assign((size_t)nf * nlegmax, 0); histo.assign((size_t)nf * (maxhisto + 1), 0);
        dens.assign((size_t)T.nstates * T.nstates, 0);
        outer_hist.assign(T.nsst, 0); sign_hist.assign(T.nsst, 0); contrib_sst.assign(T.nsst, 0);
        outerstate_hist.assign(T.nstates, 0); signstate_hist.assign(T.nstates, 0); contrib_state.assign(T.nstates, 0);
        accpair.assign((size_t)nf * ntau, 0);
        giw.assign((size_t)2 * nf * 2 * ngiw, 0.0);
        gtau_full.assign(offdiag ? (size_t)nf * nf * ntau : 0, 0.0);
        meas_sign = 0; cnt_meas = 0;
    }
};

}  // namespace

// ====================================================================== C interface (ctypes)
extern "C" {

void* orc_create(int nsst, int nb, const int* dim, const int* offset, const int* connect,
                 const long long* op_off, const double* opdata, long long nopdata,
                 const double* evals, double egs, const int* first_state, const int* state2sst,
                 const double* ftau_diag, int nftau,
                 double beta, int ntau, int nlegmax, int maxhisto, int ncorr,
                 double pglob, double pshift, double p4op, double pouter, double taudiffmax_cfg,
                 int has_szt, int meas_dm) {
    Chain* c = new Chain();
    Tables& T = c->T;
    T.nsst = nsst; T.nb = nb; T.nops = nb * 4;
    T.dim.assign(dim, dim + nsst); T.offset.assign(offset, offset + nsst);
    T.connect.assign(connect, connect + (size_t)nsst * T.nops);
    T.op_off.assign(op_off, op_off + (size_t)nsst * T.nops);
    T.opdata.assign(opdata, opdata + nopdata);
    T.nstates = 1; for (int i = 0; i < nb; ++i) T.nstates *= 4;
    T.evals.assign(evals, evals + T.nstates); T.egs = egs;
    T.first_state.assign(first_state, first_state + nsst);
    T.state2sst.assign(state2sst, state2sst + T.nstates);
    T.dmax = *std::max_element(T.dim.begin(), T.dim.end());
    c->nb = nb; c->nf = 2 * nb; c->nftau = nftau; c->beta = beta; c->ntau = ntau; c->nlegmax = nlegmax;
    c->maxhisto = maxhisto; c->ncorr = ncorr; c->pglob = pglob; c->pshift = pshift; c->p4op = p4op; c->pouter = pouter;
    c->has_szt = has_szt != 0; c->meas_dm = meas_dm;
    c->ftau.assign(ftau_diag, ftau_diag + (size_t)nb * 2 * nftau);
    c->init_window(taudiffmax_cfg);
    c->nslide = std::max(2 * nb, (int)std::floor((double)ncorr / (double)c->num_windows));   // CTQMC.F90:8069
    c->alloc_accumulators();
    c->ph = Philox{0, 0, 0, {0, 0, 0, 0}, ~0ull};
    return c;
}
void orc_destroy(void* h) { delete (Chain*)h; }

// kind 0: reference mt19937 (libpath = oracle/_ref/libmtrng.so); kind 1: Philox(seed, chain)
int orc_init_rnd(void* h, int kind, unsigned seed, unsigned chain, const char* libpath) {
    Chain* c = (Chain*)h; c->rng_kind = kind;
    if (kind == 0) {
        MtLib& m = c->mt;
        if (!m.h) {
            m.h = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
            if (!m.h) return -1;
            m.fnew = (void* (*)())dlsym(m.h, "new_random");
            m.fseed = (void (*)(void*, unsigned))dlsym(m.h, "seed_random");
            m.frand = (double (*)(void*))dlsym(m.h, "get_random");
            m.frandint = (int (*)(void*, int, int))dlsym(m.h, "get_randint");
            if (!m.fnew || !m.fseed || !m.frand || !m.frandint) return -2;
            m.rng = m.fnew();
        }
        m.fseed(m.rng, seed);
    } else {
        c->ph = Philox{seed, chain, 0, {0, 0, 0, 0}, ~0ull};
    }
    return 0;
}
double orc_grnd(void* h) { return ((Chain*)h)->grnd(); }
int orc_randint(void* h, int lo, int hi) { return ((Chain*)h)->randint(lo, hi); }
unsigned long long orc_rng_counter(void* h) { return ((Chain*)h)->ph.ctr; }

void orc_init_trace(void* h) { ((Chain*)h)->init_trace(); }
// offdiag = 1 (oracle-only groundwork): full hybridisation function F(f_a, f_c, i), flavour f = 2 b + s; before orc_init_trace
void orc_set_offdiag(void* h, const double* ftau_full) {
    Chain* c = (Chain*)h; c->offdiag = true;
    c->ftau_full.assign(ftau_full, ftau_full + (size_t)c->nf * c->nf * c->nftau);
    c->gtau_full.assign((size_t)c->nf * c->nf * c->ntau, 0.0);
}
void orc_get_gtau_full(void* h, double* out) { Chain* c = (Chain*)h; std::memcpy(out, c->gtau_full.data(), c->gtau_full.size() * sizeof(double)); }
void orc_set_meas_giw(void* h, int ngiw) { Chain* c = (Chain*)h; c->ngiw = ngiw; c->giw.assign((size_t)2 * c->nf * 2 * ngiw, 0.0); }
void orc_get_giw(void* h, double* out) { Chain* c = (Chain*)h; std::memcpy(out, c->giw.data(), c->giw.size() * sizeof(double)); }
void orc_set_statesampling(void* h, int on) { ((Chain*)h)->statesampling = on != 0; }     // before orc_init_trace
int orc_outer_state(void* h) { return ((Chain*)h)->outer_state; }
void orc_set_taudiffmax(void* h, double t) { ((Chain*)h)->init_window(t); }
// SymMoveNN lists (1-based target indices b + nb s + 1); the inverses are appended as init_Trace does (Trace.F90:471-478)
void orc_set_symmoves(void* h, int nsym, const int* moves) {
    Chain* c = (Chain*)h; const int nf = c->nf;
    c->nsym = nsym; c->symmoves.assign((size_t)2 * nsym * nf, 0);
    for (int i = 0; i < nsym; ++i) for (int j = 0; j < nf; ++j) {
        const int v = moves[i * nf + j];
        c->symmoves[(size_t)i * nf + j] = v;
        c->symmoves[(size_t)(nsym + i) * nf + v - 1] = j + 1;
    }
}
void orc_set_fast_trace(void* h, int on) { Chain* c = (Chain*)h; c->fast_trace = on; c->ec.valid = false; c->fc.valid = false; }
void orc_init_counters(void* h) { ((Chain*)h)->reset_counters(); }
void orc_set_log(void* h, double* buf, long long cap) { Chain* c = (Chain*)h; c->logp = buf; c->log_cap = cap; c->log_n = 0; }
long long orc_log_count(void* h) { return ((Chain*)h)->log_n; }
int orc_log_width() { return LOGW; }

void orc_steps(void* h, long long n) { Chain* c = (Chain*)h; for (long long i = 0; i < n; ++i) c->step(); }
void orc_warmup(void* h, long long n) { Chain* c = (Chain*)h; c->reset_counters(); for (long long i = 0; i < n; ++i) c->step(); }
void orc_measure(void* h, long long nmeas, int reset) {
    Chain* c = (Chain*)h;
    if (reset) { c->alloc_accumulators(); }
    c->reset_counters();
    for (long long im = 0; im < nmeas; ++im) { for (int ic = 0; ic < c->ncorr; ++ic) c->step(); c->measure(); }
}
void orc_measure_once(void* h) { ((Chain*)h)->measure(); }
int orc_error(void* h) { return ((Chain*)h)->error; }

// state getters
int orc_noper(void* h) { return (int)((Chain*)h)->ops.size(); }
int orc_outer_sst(void* h) { return ((Chain*)h)->outer_sst; }
int orc_window_position(void* h) { return ((Chain*)h)->window_position; }
int orc_num_windows(void* h) { return ((Chain*)h)->num_windows; }
int orc_nslide(void* h) { return ((Chain*)h)->nslide; }
void orc_get_config(void* h, double* taus, int* orbs, int* spins, int* cas) {
    Chain* c = (Chain*)h;
    for (size_t i = 0; i < c->ops.size(); ++i) { taus[i] = c->ops[i].tau; orbs[i] = c->ops[i].b + 1; spins[i] = c->ops[i].s + 1; cas[i] = c->ops[i].ca; }
}
void orc_set_config(void* h, int n, const double* taus, const int* orbs, const int* spins, const int* cas, int outer_sst, int outer_state) {
    ((Chain*)h)->set_config(n, taus, orbs, spins, cas, outer_sst, outer_state);
}
void orc_get_scalars(void* h, double* out) {   // trace log/sign, det log/sign, cfgsign, meas_sign, cnt_meas, cnt_steps
    Chain* c = (Chain*)h;
    out[0] = c->Trace.log; out[1] = c->Trace.sign; out[2] = c->Det.log; out[3] = c->Det.sign; out[4] = c->cfgsign;
    out[5] = c->meas_sign; out[6] = (double)c->cnt_meas; out[7] = (double)c->cnt_steps;
}
void orc_get_counters(void* h, double* out) {
    Counters& C = ((Chain*)h)->C;
    double v[15] = {C.TryAdd, C.TryRem, C.TryGlob, C.TryShift, C.AccAdd, C.AccRem, C.AccGlob, C.AccShift, C.AccQNAdd, C.AccQNRem, C.AccQNGlob,
                    C.TryAdd4, C.AccAdd4, C.TryRem4, C.AccRem4};
    std::memcpy(out, v, sizeof(v));
}
int orc_kf(void* h, int f) { return ((Chain*)h)->kf[f]; }
void orc_get_minv(void* h, int f, double* out) { Chain* c = (Chain*)h; std::memcpy(out, c->M[f].data(), c->M[f].size() * sizeof(double)); }
// from-scratch inverse / determinant of the current configuration (pin for the fast updates)
void orc_scratch_det(void* h, double* logdet, double* sign) {
    Chain* c = (Chain*)h; std::vector<std::vector<double>> Mn; std::vector<int> kn;
    LogVal d = c->build_inverse_all(c->ops, Mn, kn); *logdet = d.log; *sign = d.sign;
}
void orc_scratch_minv(void* h, int f, double* out) {
    Chain* c = (Chain*)h; std::vector<std::vector<double>> Mn; std::vector<int> kn;
    c->build_inverse_all(c->ops, Mn, kn); std::memcpy(out, Mn[f].data(), Mn[f].size() * sizeof(double));
}
double orc_trace_of(void* h, int outer_sst, double* sign) {
    Chain* c = (Chain*)h; LogVal t = c->trace_full(c->ops, outer_sst, nullptr); *sign = t.sign; return t.log;
}
int orc_npairs(void* h) { Chain* c = (Chain*)h; return c->count_pairs(c->ops); }
void orc_shift_window(void* h) { ((Chain*)h)->shift_window(); }

// accumulators (raw sums; normalisation CTQMC.F90:8663-8771 is done in oracle/results.py)
void orc_get_accum(void* h, double* gtau, double* gleg, double* histo, double* dens, double* outer_hist,
                   double* sign_hist, double* contrib_sst, double* outerstate_hist, double* signstate_hist,
                   double* contrib_state, long long* accpair) {
    Chain* c = (Chain*)h;
    auto cp = [](double* d, const std::vector<double>& s) { if (d) std::memcpy(d, s.data(), s.size() * sizeof(double)); };
    cp(gtau, c->gtau); cp(gleg, c->gleg); cp(histo, c->histo); cp(dens, c->dens); cp(outer_hist, c->outer_hist);
    cp(sign_hist, c->sign_hist); cp(contrib_sst, c->contrib_sst); cp(outerstate_hist, c->outerstate_hist);
    cp(signstate_hist, c->signstate_hist); cp(contrib_state, c->contrib_state);
    if (accpair) std::memcpy(accpair, c->accpair.data(), c->accpair.size() * sizeof(long long));
}

}  // extern "C"
