// CT-HYB chain engine: one warp advances one Markov chain.
//
// This file holds the per-chain logic of the hot path; it is compiled
//  * by nvcc for sm_100a (32 lanes per chain, kernels in engine.hpp), and
//  * by g++ with W2B_HOST_EMU (1 lane) into a test-only library for CPU logic tests.
//
// Reference semantics restated here (files under /root/reference/src/ctqmc_fortran):
//   scheduler + window shift        CTQMC.F90:8213-8258, 8428-8533 ; Trace.F90:1608-1751
//   StepAdd / StepRem               CTQMC.F90:3353-3612, 3892-4150
//   pair_OperAdd / process_OperAdd  Trace.F90:3657-3698, 4275-4462   (FPos = rank among same flavour & type)
//   gen_OperRemove / NPairs         Trace.F90:4851-4950, 1770-1935
//   check_sst_sequence              Trace.F90:1243-1302
//   get_LogDetRatPairAdd/Rem        Trace.F90:2775-2832, 2937-2963 ; MatrixUpdate.F90:392-482
//   get_Trace_EB (scalar blocks)    Trace.F90:5556-5727
//   get_Sign                        Trace.F90:3479-3618
//   StepGlob / StepShiftTau         CTQMC.F90:4433-4835 ; Trace.F90:1103-1433, 2126-2188, 2279-2457
//   measurements                    CTQMC.F90:2252-2440, 3088-3151, 3264-3350
//
// B200-first representation (NOT the reference's linked lists / per-move heap churn):
//   * operators: one time-sorted packed array per chain (tau f64, code u8 = flavour*2+type, sst i16 = superstate
//     after the operator); FPos, NPairs, window membership are index arithmetic / lane-parallel counts;
//   * inverse hybridisation matrices M_f: chain-major, row-major [creator][annihilator] with leading dimension ld,
//     updated out-of-place into one spare buffer per chain (rank-1 insert / remove with the row/column shift fused
//     into the same pass), reduced with warp shuffles;
//   * local trace for 1x1 blocks: the ratio is evaluated only over the changed sub-string (exact, not cached);
//   * local trace for larger blocks: cached ket / bra partial products, evaluated as 8-column jobs on the FP64 tensor
//     pipe by a kernel of its own (trace_job below, trace_matrix.hpp);
//   * two schedulers over the chain population: a persistent CTA phase scheduler (1x1 blocks) and rounds of
//     propose / trace / decide kernels with global moves on a second stream (blocked traces);
//   * counter-based RNG (Philox4x32-10) consumed in the reference's draw order.
#pragma once
#include "warp.hpp"
#include "trace_matrix.hpp"

namespace w2b {

constexpr int LOGW = 20;      // doubles per replay-log record (same layout as oracle/chain.cpp)
constexpr int NF_MAX = 14;
enum { C_TRYADD = 0, C_TRYREM, C_TRYGLOB, C_TRYSHIFT, C_ACCADD, C_ACCREM, C_ACCGLOB, C_ACCSHIFT, C_QNADD, C_QNREM, C_QNGLOB, C_STEPS, C_ALGBYTES, C_TRFLOPS, C_TRYADD4, C_ACCADD4, C_TRYREM4, C_ACCREM4, C_FOUR_RANK2, C_NCNT };
enum { IH_NOPS = 0, IH_OUTER, IH_WPOS, IH_NPAIRS, IH_PERM, IH_PALT, IH_ERR, IH_KV, IH_KF = 8, IH_MSLOT = IH_KF + NF_MAX, IH_BV = IH_MSLOT + NF_MAX + 1, IH_NFREE, IH_OSTATE, IH_STRIDE };   // IH_OSTATE: outer eigenstate inside the outer superstate (state sampling; else 0)
enum { DH_DETLOG = 0, DH_DETSIGN, DH_TRLOG, DH_TRSIGN, DH_CFGSIGN, DH_STRIDE = 8 };
enum { ERR_CAPACITY = 1, ERR_DETRAT_SMALL = 2, ERR_PAIR_LOGIC = 4, ERR_GLOBAL_QN = 8, ERR_SINGULAR = 16 };

// static tables (device pointers), identical for all chains
struct Tables {
    int nb, nf, nsst, nops, dmax, nftau, nstates;
    double beta, egs;
    const short* connect;        // [nsst*nops]
    const double* e0;            // [nsst] lowest eigenvalue of each block
    const double* elem;          // [nsst*nops] operator element for 1x1 blocks (0 where not connected)
    const double* lelem;         // [nsst*nops] log|element| (0 where not connected)
    const double* ftau;          // [nf][nftau] diagonal hybridisation, flavour f = 2*band+spin
    const int* first_state;      // [nsst]
    const int* state2sst;        // [nstates]
    const int* dim;              // [nsst]
    const int* offset;           // [nsst]
    const int* dens_off;         // [nsst] offset of the block's d x d slab inside the density-matrix accumulator
    const double* empty_w;       // [nsst] normalised empty-trace weights (init_Trace / StepChangeOuter)
    const double* empty_log;     // [nsst] log of the empty trace of each outer superstate
    double empty_norm;           // sum over ALL states of exp(-(E - Egs) beta): normalisation of the per-state empty-trace weights (state sampling)
    const double* leg_a; const double* leg_b;   // Legendre recurrence coefficients (2l-1)/l and (l-1)/l, [NLegMax]
    const int* symmoves;         // [2 NSymMove][nf] flavour permutations (target index b + nb s + 1) and their inverses (Trace.F90:471-478)
    const double* leg_g; const double* leg_c;   // monic recurrence g_l = (l-1)^2/((2l-1)(2l-3)) and leading coefficients c_l: P_l = c_l Q_l
    const OpDesc* desc;          // [nsst*nops] packed operator records of the blocked trace
    const long long* op_off;     // matrix path
    const double* opdata;
    const double* opfrag;        // operator blocks in DMMA A-fragment order (OpDesc::frag_off)
    const double* evals;
};

struct Run {
    double taudiff_max, pglob, pshift, p4op, pouter;
    int num_windows, nslide, ntau, nlegmax, maxhisto, has_szt, nbins, meas_dm;
    unsigned seed, chain_base;
    int naccpair; double accpair_max;   // accepted-pair-distance histogram: bins / range (Ntau, beta; calibration phase 2: grid, taudiff_max)
    int freeze_window;           // ctqmc_calibrate never shifts the window (CTQMC.F90:8118-8175 has no shift_window call)
    int propose_run;             // consecutive early-exit moves one chain may take inside one propose phase
    int prefetch;                // software prefetch of the next work item's records (0 = off)
    int same_key;                // diagnostics only: every chain uses chain 0's RNG key (identical control flow across warps)
    int nsym;                    // user symmetry moves (NSymMove); the table Tables::symmoves holds them and their inverses
    int ngiw;                    // MeasGiw != 0: Matsubara frequencies n = -ngiw .. ngiw - 1 of the G(iw) estimator (Niw); 0 = not measured
    int four_rank2;              // accepted 4-operator moves: rank-2 update of the stored inverses (1, default) or re-inversion from scratch as the reference does (0)
    int statesampling;           // statesampling = 1: the local trace runs over ONE outer eigenstate (Trace.F90:203, 1477-1604)
    unsigned stream;             // solves since init_rnd: every init_solver draws from a fresh section of each chain's Philox stream
};
constexpr int STREAM_SHIFT = 44;     // draws per chain and solve before two solves could overlap: 2^44

// a proposal that survived the quantum-number check travels from the propose kernel to the decide kernel in this record
struct Prop {
    double t1, t2, tau_a, tau_c, factor, dlog, dsign;
    int f, i1, i2, p1, p2, begin_sst, nother, type1;
    int new_state;                           // outer pair moves with state sampling: chosen outer eigenstate (its WeightedOuterStateChoice factor is folded into `factor`)
};

// a global move / tau shift between its two halves (prepare -> trace kernel -> finish)
struct RareJob {
    double dt, factor, dn_log, dn_sign;
    int kind, new_outer, type, stored;       // kind 3 = global move, 4 = tau shift; stored: the trace kernel filled the ket caches of the new configuration
    int knew[NF_MAX];
    double tn_log, tn_sign;                  // staged global moves of the CTA scheduler: local trace of the proposed configuration
    int nnew;                                // kinds 5 - 8 (outer pair moves, 4-operator moves): number of operators of the proposed sequence in the scratch arrays
    int np_new;                              // kinds 7 / 8: NPairs of the proposed configuration
    int new_state;                           // state sampling: outer eigenstate of the proposed configuration (-1: sum over the superstate)
};

// round scheduler counters (device memory, two sets used in alternation by consecutive rounds)
enum { SC_NQ0 = 0, SC_NQ1, SC_NQ2, SC_TICKET, SC_NRUN, SC_MAXLEFT, SC_TTICKET, SC_NT0, SC_NT1, SC_NT2, SC_NRFIN, SC_STRIDE = 16 };
// per-chain scheduling state (Chains::parked): global moves / tau shifts leave the round pipeline and come back through these states
enum { ST_RUN = 0, ST_RARE_NEW = 1, ST_RARE_TRACE = 3, ST_RARE_QUEUED = 4, ST_RARE_DONE = 5 };
constexpr int Q_MEAS = 1 << 28;      // trace-queue entry of a measurement: extend both cache frontiers to tau = beta/2, no trace
constexpr int Q_RARE = 1 << 29;      // trace-queue entry of a global move / tau shift (whole sequence, no caches)
constexpr int Q_REM = 1 << 30;       // decide-queue entry = chain id | Q_REM for a removal
constexpr int TRACE_COLS = TILE_COLS;    // outer states (columns) per trace job = one DMMA column tile
constexpr int Q_NTSHIFT = 26;         // trace-queue entry: + column tile << Q_NTSHIFT
constexpr int Q_IDMASK = (1 << 26) - 1;

struct Chains {
    int nchains, ops_cap, kcap, ld;
    double* op_tau; unsigned char* op_code; short* op_sst;
    double* tau_c; double* tau_a;
    double* M;
    int* ih; double* dh;
    unsigned long long* rng_ctr; unsigned long long* counters;
    double* sc_tau; unsigned char* sc_code; short* sc_sst;     // per-chain scratch (global / tau-shift moves)
    double* sc_tc; double* sc_ta;                              // [chain][nf][kcap]
    // blocked-trace caches (dmax > 1 only): slot ids per operator, free-slot stack, ket / bra matrices + column log scales
    unsigned short* op_slot; unsigned short* free_slot;        // [chain][ops_cap]
    double* ket; double* bra;                                  // [chain][ops_cap][mstride]   (compact d x m, column-major)
    double* ketlg; double* bralg;                              // [chain][ops_cap][dmax]
    double* pt;                                                // [chain][2*dmax] (log, sign) of every outer state's trace contribution
    int mstride, dmax;
    // round scheduler: moves left in the current sweep, rare-move flags, proposals, work queues, counters
    int* left; unsigned char* parked; unsigned char* kind; Prop* prop; RareJob* rjob;
    int* q_rfin; int* q_dec0; int* q_dec1; int* q_dec2; int* sched;      // q_rare, q_rfin and the trace queues: one buffer per counter set
    double* ptc;                                               // trace kernel results: per-outer-state parts (log, sign) of the proposed configuration [chain][2*dmax]
    int* q_tr0; int* q_tr1; int* q_tr2;                        // trace jobs (chain | column tile | removal flag) by number of steps: long, medium, short
    int ss;                                                    // state sampling: trace jobs run for the column tile of the sampled outer eigenstate only
    int meas_pre;                                              // measurement kernel: the cache frontiers already reach beta/2 (extended by trace jobs)
    int ldx, xdoubles;                                         // work matrix: leading dimension (4 mod 16) and size ldx * dmax
};

// accumulators (device), one slab per bin; offsets in doubles
struct Accum {
    double* base; long long stride;
    int o_gtau, o_gleg, o_histo, o_dens, o_outer, o_signsst, o_contribsst, o_outerst, o_signst, o_contribst, o_sign, o_cnt;
    int o_giw;                       // [nf][2 ngiw] complex (re, im) sums of the Matsubara estimator
    unsigned long long* accpair;     // [nf][ntau]
    int ndens;
    double* pairlist; int* pairlist_cnt; int pairlist_cap;      // AccPairTau of ctqmc_calibrate phase 1 (CTQMC.F90:3583-3586, 4123-4126): tau distances of the first accepted pair moves
};

struct Ctx {
    double* tau; unsigned char* code; short* sst; double* tc; double* ta; double* Mbase; int* ih; double* dh;
    unsigned long long* cnt;
    double* sc_tau; unsigned char* sc_code; short* sc_sst; double* sc_tc; double* sc_ta;
    double* wQ; double* wR; double* wV; int* wI;
    unsigned short* slot; unsigned short* freel; double* ket; double* bra; double* ketlg; double* bralg; double* pt;
    double* mX; double* mY; double* mlg; double* mpt; double* mpw;   // per-warp work space of the blocked trace (mY: measurement kernel only)
    int ldx;
    int kv, bv, nfree;
    int n, outer, wpos, npairs, perm, palt, err;
    int ostate, only;      // outer eigenstate (0 without state sampling); only = ostate with state sampling, else -1 (argument of parts_sum)
    double det_log, det_sign, tr_log, tr_sign, cfgsign, wmin, wmax;
    Philox rng;
    double* log;           // replay record for the current step (or nullptr)
};

#if W2B_DEVICE
W2B_HD double sgn1(double x) { return __double2hiint(x) < 0 ? -1.0 : 1.0; }      // Fortran sign(1.0, x): the sign bit (also of -0.0)
#else
W2B_HD double sgn1(double x) { return std::signbit(x) ? -1.0 : 1.0; }
#endif
W2B_HD int conn(const Tables& T, int sst, int opc) { return (int)ldg(&T.connect[sst * T.nops + opc]); }

// get_HybrF Trace.F90:1014-1033
W2B_HDN double hybrF(const Tables& T, double tau, int f) {
    double sig = 1.0;
    if (tau < 0.0) { tau += T.beta; sig = -1.0; }
    double n = tau / T.beta * (double)(T.nftau - 1);
    int i = (int)n;
    const double* p = T.ftau + (size_t)f * T.nftau + i;
    double a = ldg(p), b = ldg(p + 1);
    return sig * (a + (n - (double)i) * (b - a));
}

W2B_HD void set_window_edges(const Tables& T, const Run& R, Ctx& x) {   // set_window Trace.F90:1697-1699
    if (R.num_windows > 1) {
        x.wmin = ((double)x.wpos / (double)(R.num_windows + 1)) * T.beta;
        x.wmax = ((double)(x.wpos + 2) / (double)(R.num_windows + 1)) * T.beta;
    } else { x.wmin = 0.0; x.wmax = T.beta; }
}

W2B_HD double* Mbuf(const Chains& C, Ctx& x, int f) { return x.Mbase + (size_t)x.ih[IH_MSLOT + f] * ((size_t)C.kcap * C.ld); }

// number of sorted entries < t ; eq is set when an entry equals t
W2B_HDN int count_lt(const double* a, int n, double t, bool& eq) {
    int c = 0; bool e = false;
#pragma unroll 1
    for (int i = lane_id(); i < n; i += W2B_WARP) { double v = a[i]; c += (v < t); e |= (v == t); }
    eq = wany(e);
    return wsum(c);
}
W2B_HD int count_lt(const double* a, int n, double t) { bool e; return count_lt(a, n, t, e); }
// both insertion positions in one pass over the operator times (t1 < t2); eq: some entry equals t1 or t2
W2B_HDN void count_lt2(const double* a, int n, double t1, double t2, int& c1, int& c2, bool& eq) {
    int x1 = 0, x2 = 0; bool e = false;
#pragma unroll 1
    for (int i = lane_id(); i < n; i += W2B_WARP) { double v = a[i]; x1 += (v < t1); x2 += (v < t2); e |= (v == t1) | (v == t2); }
    eq = wany(e);
    const int packed = wsum(x1 | (x2 << 16));                 // n <= 4096 operators: both counts fit 16 bits
    c1 = packed & 0xffff; c2 = packed >> 16;
}
W2B_HDN int count_le(const double* a, int n, double t) {
    int c = 0;
#pragma unroll 1
    for (int i = lane_id(); i < n; i += W2B_WARP) c += (a[i] <= t);
    return wsum(c);
}

// array shifts by whole-warp chunks (read chunk, sync, write)
template <class T_> W2B_HDN void shift_up(T_* a, int from, int n, int by) {       // [from,n) -> [from+by, n+by)
#pragma unroll 1
    for (int top = n; top > from; top -= W2B_WARP) {
        int idx = top - 1 - lane_id();
        T_ v = T_(); bool act = idx >= from;
        if (act) v = a[idx];
        wsync();
        if (act) a[idx + by] = v;
        wsync();
    }
}
template <class T_> W2B_HDN void shift_down(T_* a, int from, int n, int by) {     // [from+by, n) -> [from, n-by)
#pragma unroll 1
    for (int base = from; base < n - by; base += W2B_WARP) {
        int idx = base + lane_id();
        T_ v = T_(); bool act = idx < n - by;
        if (act) v = a[idx + by];
        wsync();
        if (act) a[idx] = v;
        wsync();
    }
}

// Serial superstate walk over `m` existing operators starting at list index i0 (codes from `code`), starting from
// superstate `cur`; stores the superstate after each operator into out[i0out + j].  Returns the final superstate
// (-1 if annihilated on the way).
W2B_HDN int walk_store(const Tables& T, const unsigned char* code, int i0, int m, int cur, short* out, int i0out) {
#pragma unroll 1
    for (int base = 0; base < m; base += W2B_WARP) {
        int j = base + lane_id(); bool act = j < m;
        int cj = act ? (int)code[i0 + j] : 0;
        int cnt = m - base < W2B_WARP ? m - base : W2B_WARP;
        int mine = -1;
#pragma unroll 1
        for (int q = 0; q < cnt; ++q) {
            int cq = wbcast(cj, q);
            cur = cur >= 0 ? conn(T, cur, cq) : -1;
            if (lane_id() == q) mine = cur;
        }
        if (act) out[i0out + j] = (short)mine;
    }
    wsync();
    return cur;
}

// Re-evaluate `m` existing operators [i0, i0+m) from new superstate `cur` while the old superstates are read from
// x.sst; accumulates the change of the scalar local trace (Trace.F90:5595-5642 restricted to the changed sub-string).
// On entry: old_before = old superstate before operator i0, tl = time of the left end of the first interval.
// On exit:  cur / old_after = new / old superstate after the last walked operator, tl = its time (unchanged if m == 0),
//           nother = number of walked operators whose flavour differs from f.
W2B_HDN void walk_delta(const Tables& T, const Ctx& x, int i0, int m, int f, int& cur, int& old_after, double& tl,
                       double& dlog, double& dsign, int& nother) {
    double acc = 0.0; int neg = 0, oth = 0;
    int old_before0 = old_after; double tl0 = tl;
#pragma unroll 1
    for (int base = 0; base < m; base += W2B_WARP) {
        int j = base + lane_id(); bool act = j < m;
        int cj = act ? (int)x.code[i0 + j] : 0;
        int cnt = m - base < W2B_WARP ? m - base : W2B_WARP;
        int nb4 = -1;
#pragma unroll 1
        for (int q = 0; q < cnt; ++q) {
            int cq = wbcast(cj, q);
            if (lane_id() == q) nb4 = cur;
            cur = cur >= 0 ? conn(T, cur, cq) : -1;
        }
        if (act && nb4 >= 0) {
            double tj = x.tau[i0 + j];
            int ob = (j == 0) ? old_before0 : (int)x.sst[i0 + j - 1];
            double tp = (j == 0) ? tl0 : x.tau[i0 + j - 1];
            acc += -(ldg(&T.e0[nb4]) - ldg(&T.e0[ob])) * (tj - tp);
            double en = ldg(&T.elem[nb4 * T.nops + cj]), eo = ldg(&T.elem[ob * T.nops + cj]);
            if (en != 0.0) { acc += ldg(&T.lelem[nb4 * T.nops + cj]) - ldg(&T.lelem[ob * T.nops + cj]); neg += (en < 0.0) != (eo < 0.0); }
            oth += ((cj >> 1) != f);
        }
    }
    if (m > 0) { old_after = (int)x.sst[i0 + m - 1]; tl = x.tau[i0 + m - 1]; }
    dlog += wsum(acc);
    if (wsum(neg) & 1) dsign = -dsign;
    nother = wsum(oth);
}

// pairs (Trace.F90:1794-1814): earlier >= wmin, later <= wmax, !(later > earlier + taudiff_max)
W2B_HD bool pairable(const Run& R, const Ctx& x, double a, double b) {
    double e = a < b ? a : b, l = a < b ? b : a;
    return e >= x.wmin && l <= x.wmax && !(l > e + R.taudiff_max);
}
W2B_HDN int pairs_with(const Run& R, const Ctx& x, const double* opp, int k, double t) {
    int c = 0;
#pragma unroll 1
    for (int i = lane_id(); i < k; i += W2B_WARP) c += pairable(R, x, opp[i], t);
    return wsum(c);
}
// The annihilator times are sorted: only those within taudiff_max of the creator can pair, so the exact predicate is evaluated on
// the candidates between two binary searches (bounds widened by a relative 1e-12: NPairs is integer state, the predicate decides).
// With 65 windows (C5) the window moves every second attempt and every move recounts: the full k x k scan was 14 % of the sweep
// kernel's instructions.
W2B_HDN int count_pairs_full(const Tables& T, const Run& R, const Chains& C, const Ctx& x) {
    int c = 0;
    const double slack = 1e-12 * T.beta;
#pragma unroll 1
    for (int f = 0; f < T.nf; ++f) {
        int k = x.ih[IH_KF + f];
        const double* tc = x.tc + (size_t)f * C.kcap; const double* ta = x.ta + (size_t)f * C.kcap;
#pragma unroll 1
        for (int i = lane_id(); i < k; i += W2B_WARP) {
            double t = tc[i];
            if (t < x.wmin || t > x.wmax) continue;
            const double lo_t = t - R.taudiff_max - slack, hi_t = t + R.taudiff_max + slack;
            int lo = 0, hi = k;
#pragma unroll 1
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (ta[mid] < lo_t) lo = mid + 1; else hi = mid; }
#pragma unroll 1
            for (int a = lo; a < k; ++a) { const double v = ta[a]; if (v > hi_t) break; c += pairable(R, x, v, t); }
        }
    }
    return wsum(c);
}

// number of creators sitting at an even position of the flavour's time-ordered list (get_permaltca, Trace.F90:3554-3588)
W2B_HDN int palt_flavour(const Chains& C, const Ctx& x, int f) {
    int k = x.ih[IH_KF + f];
    const double* tc = x.tc + (size_t)f * C.kcap; const double* ta = x.ta + (size_t)f * C.kcap;
    int c = 0;
#pragma unroll 1
    for (int i = lane_id(); i < k; i += W2B_WARP) {
        const double t = tc[i];
        int lo = 0, hi = k;                                   // annihilators before this creator: lower bound in the sorted times
#pragma unroll 1
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (ta[mid] < t) lo = mid + 1; else hi = mid; }
        c += (((i + lo) & 1) == 0);
    }
    return wsum(c) & 1;
}
// inversion parity of the flavour sequence in time order (== get_BlockSign, Trace.F90:3594-3618)
W2B_HDN int perm_full(const unsigned char* code, int n) {
    int c = 0;
#pragma unroll 1
    for (int i = lane_id(); i < n; i += W2B_WARP) {
        int fi = code[i] >> 1;
#pragma unroll 1
        for (int j = i + 1; j < n; ++j) c += ((code[j] >> 1) < fi);
    }
    return wsum(c) & 1;
}
W2B_HDN void update_cfgsign(const Tables& T, Ctx& x) {
    int pal = 0;
#pragma unroll 1
    for (int f = 0; f < T.nf; ++f) pal ^= (x.palt >> f) & 1;
    if ((x.n / 2) & 1) pal ^= 1;
    double s = x.tr_sign * x.det_sign * ((x.perm & 1) ? -1.0 : 1.0) * (pal ? -1.0 : 1.0);
    x.cfgsign = -s;                                         // CTQMC.F90:3574
}

// ------------------------------------------------------------------ blocked local trace (dmax > 1): cache frontiers
// (design notes in trace_matrix.hpp; reference get_Trace_EB Trace.F90:5354-6033, update_trace_EB 6109-6209)
W2B_HD int sst_before(const Ctx& x, int i) { return i == 0 ? x.outer : (int)x.sst[i - 1]; }
W2B_HD double* ket_slot(const Chains& C, const Ctx& x, int i) { return x.ket + (size_t)x.slot[i] * C.mstride; }
W2B_HD double* bra_slot(const Chains& C, const Ctx& x, int i) { return x.bra + (size_t)x.slot[i] * C.mstride; }

// all cache slots back to "operator i owns slot i", nothing valid
W2B_HD void slots_reset(const Chains& C, Ctx& x) {
#pragma unroll 1
    for (int i = lane_id(); i < C.ops_cap; i += W2B_WARP) { x.slot[i] = (unsigned short)i; x.freel[i] = (unsigned short)(C.ops_cap - 1 - i); }
    x.nfree = C.ops_cap - x.n; x.kv = 0; x.bv = x.n;
    wsync();
}

// make KET_i valid for all i <= upto; returns true when the work matrix holds KET_upto (with its scales in x.mlg) on exit
W2B_HDN bool extend_kets(const Tables& T, const Chains& C, Ctx& x, int upto) {
    if (x.kv > upto) return false;
    const int m = ldg(&T.dim[x.outer]), ld = x.ldx;
    int i = x.kv, sst; double tprev;
    double* X = x.mX;
    if (i == 0) { mat_identity(X, ld, x.mlg, m); sst = x.outer; tprev = 0.0; }
    else {
        sst = (int)x.sst[i - 1]; tprev = x.tau[i - 1];
        const int d = ldg(&T.dim[sst]);
        mat_load_padded(X, ld, ket_slot(C, x, i - 1), d, m);
        mat_copy(x.mlg, x.ketlg + (size_t)x.slot[i - 1] * C.dmax, m);
        wsync();
    }
#pragma unroll 1
    for (; i <= upto; ++i) {
        sst = mat_apply(T, sst, (int)x.code[i], x.tau[i] - tprev, X, ld, x.mlg, x.mpw, m);
        const int d = ldg(&T.dim[sst]);
        mat_move(ket_slot(C, x, i), d, X, ld, d, m);
        mat_copy(x.ketlg + (size_t)x.slot[i] * C.dmax, x.mlg, m);
        tprev = x.tau[i];
    }
    wsync();
    x.kv = upto + 1;
    return true;
}

// make BRA_i valid for all i >= downto
W2B_HDN void extend_bras(const Tables& T, const Chains& C, Ctx& x, int downto) {
    if (x.bv <= downto) return;
    const int m = ldg(&T.dim[x.outer]), ld = x.ldx;
    int i = x.bv; double tnext;
    double* X = x.mX;
    if (i == x.n) { mat_identity(X, ld, x.mlg, m); tnext = T.beta; }
    else {
        tnext = x.tau[i];
        const int d = ldg(&T.dim[sst_before(x, i)]);
        mat_load_padded(X, ld, bra_slot(C, x, i), d, m);
        mat_copy(x.mlg, x.bralg + (size_t)x.slot[i] * C.dmax, m);
        wsync();
    }
#pragma unroll 1
    for (i = x.bv - 1; i >= downto; --i) {          // BRA_i = O_i^T P(tau_{i+1} - tau_i) BRA_{i+1}
        const int sb = mat_apply(T, (int)x.sst[i], (int)x.code[i] ^ 1, tnext - x.tau[i], X, ld, x.mlg, x.mpw, m);
        const int d = ldg(&T.dim[sb]);
        mat_move(bra_slot(C, x, i), d, X, ld, d, m);
        mat_copy(x.bralg + (size_t)x.slot[i] * C.dmax, x.mlg, m);
        tnext = x.tau[i];
    }
    wsync();
    x.bv = downto;
}

// from-scratch trace of an arbitrary sequence (global moves, tau shifts, set_mc_config); no caches touched
W2B_HDN void full_trace(const Tables& T, Ctx& x, const double* tau, const unsigned char* code, int n, int outer, double& tlog, double& tsign, int only = -1) {
    const int m = ldg(&T.dim[outer]), ld = x.ldx;
    double* X = x.mX;
    mat_identity(X, ld, x.mlg, m);
    int sst = outer; double tprev = 0.0;
#pragma unroll 1
    for (int i = 0; i < n; ++i) { sst = mat_apply(T, sst, (int)code[i], tau[i] - tprev, X, ld, x.mlg, x.mpw, m); tprev = tau[i]; }
    mat_close(T, sst, T.beta - tprev, X, ld, x.mlg, (const double*)nullptr, (const double*)nullptr, x.mpw, m, x.mpt, tlog, tsign, only);
}
// ------------------------------------------------------------------ trace kernel (blocked traces, pair moves)
// One job = 8 columns (outer states) of the local trace of one proposed pair insertion / removal.  Columns never mix
// (KET_i[:, j] = O_i P ... O_1 P e_j), so an outer block of dimension m splits into ceil(m/8) independent jobs -- four times
// the parallelism for the 28-wide blocks, a 2.3 kB work matrix per warp, and one DMMA column tile held in registers.
// A job extends the bra frontier down to the end of the changed sub-string, the ket frontier up to its start, multiplies
// through the sub-string and closes.  The three phases run as ONE loop over "steps" (operator, propagation time, input
// block, cache slot to fill) whose records are prepared 32 at a time, one per lane, and broadcast by shuffles: the
// operator product is the only code in the loop and the per-step global-memory latency is paid once per 32 steps.
// Writes the per-outer-state parts (log, sign) of the proposed configuration to C.ptc; the decide kernel sums them and
// moves the frontiers (kv = max(kv, s), bv = min(bv, e)).
// per-warp work space of a trace job in doubles: tile (ldx x 8) + 8 log scales + 32 propagation factors
#if W2B_DEVICE
__host__ __device__
#endif
inline int trace_ws_doubles(int ldx) { return ldx * TILE_COLS + TILE_COLS + MAT_DMAX; }
#if !W2B_DEVICE
struct Tile { double* X; double* lg; double* pw; int ld; };                                // host emulation: same interface on plain loops
template <class TablesT> inline int tile_apply(const TablesT& T, int sst_in, int opc, double dt, const Tile& w, int mw, double* slot, double* slot_lg, int j0, int* work) {
    return mat_apply_slot(T, sst_in, opc, dt, w.X, w.ld, w.lg, w.pw, mw, slot, slot_lg, j0, work);
}
inline void tile_load(const Tile& w, const double* src, const double* src_lg, int d, int j0, int mw) {
#pragma unroll 1
    for (int j = 0; j < mw; ++j) for (int r = 0; r < d; ++r) w.X[r + j * w.ld] = src ? src[r + j * d] : (r == j0 + j ? 1.0 : 0.0);
#pragma unroll 1
    for (int j = 0; j < mw; ++j) w.lg[j] = src ? src_lg[j] : 0.0;
}
template <class TablesT> inline void tile_close(const TablesT& T, int sst, double dt, const Tile& w, const double* B, const double* blg, int mw, int j0, double* pt) {
    mat_close_parts(T, sst, dt, w.X, w.ld, w.lg, B, blg, w.pw, mw, j0, pt);
}
#endif
// frontier tile of a chain: columns j0.. of KET_{i} (i >= 0) / BRA_{i} (i < n) or of the identity at the chain's ends
struct JobView {
    const double* tau; const unsigned char* code; const short* sst; const unsigned short* slot;
    double* ket; double* bra; double* ketlg; double* bralg;
    int n, outer, mstride, dmax;
};
W2B_HD void job_load_ket(const Tables& T, const JobView& V, const Tile& w, int i, int j0, int mw) {      // KET_i, i = -1: identity
    if (i < 0) { tile_load(w, nullptr, nullptr, ldg(&T.dim[V.outer]), j0, mw); return; }
    const int d = ldg(&T.dim[(int)V.sst[i]]);
    const size_t sl = V.slot[i];
    tile_load(w, V.ket + sl * V.mstride + j0 * d, V.ketlg + sl * V.dmax + j0, d, j0, mw);
}
W2B_HD void job_load_bra(const Tables& T, const JobView& V, const Tile& w, int i, int j0, int mw) {      // BRA_i, i = n: identity
    if (i >= V.n) { tile_load(w, nullptr, nullptr, ldg(&T.dim[V.outer]), j0, mw); return; }
    const int d = ldg(&T.dim[i == 0 ? V.outer : (int)V.sst[i - 1]]);
    const size_t sl = V.slot[i];
    tile_load(w, V.bra + sl * V.mstride + j0 * d, V.bralg + sl * V.dmax + j0, d, j0, mw);
}
// Measured alternative (round 2): ONE job per proposal over all column tiles of the outer block, sharing step records, operator
// descriptors and exp() factors: 15 % fewer instructions, but 57.9 instead of 39.7 ms per C4 step -- the kernel is bound by the
// latency of its longest jobs (a warp issues one instruction every ~16 cycles), so four short jobs on four warps beat one long one.
W2B_HD void trace_job(const Tables& T, const Chains& C, int entry, double* ws) {
    const int lane = lane_id();
    const Tile w{ws, ws + C.ldx * TILE_COLS, ws + C.ldx * TILE_COLS + TILE_COLS, C.ldx};
    const int chain = entry & Q_IDMASK, nt = (entry >> Q_NTSHIFT) & 3;
    const bool rem = (entry & Q_REM) != 0, rare = (entry & Q_RARE) != 0, meas = (entry & Q_MEAS) != 0;
    const int* ih = C.ih + (size_t)chain * IH_STRIDE;
    const size_t cb = (size_t)chain * C.ops_cap;
    JobView V;
    V.tau = C.op_tau + cb; V.code = C.op_code + cb; V.sst = C.op_sst + cb; V.slot = C.op_slot + cb;
    V.ket = C.ket + cb * C.mstride; V.bra = C.bra + cb * C.mstride; V.ketlg = C.ketlg + cb * C.dmax; V.bralg = C.bralg + cb * C.dmax;
    V.n = ih[IH_NOPS]; V.outer = ih[IH_OUTER]; V.mstride = C.mstride; V.dmax = C.dmax;
    if (rare && C.rjob[chain].kind >= 5) V.n = C.rjob[chain].nnew;       // outer pair move: the staged sequence has two operators more / less
    const int n = V.n;
    int kv = ih[IH_KV], bv = ih[IH_BV];
    const Prop& P = C.prop[chain];
    int i1 = P.i1, i2 = P.i2;
    bool store_seg = false;
    if (meas) {                                                   // kets of the operators with tau <= beta/2, bras of the others (update_b2_states_eb)
        if (n == 0 || nt * TILE_COLS >= ldg(&T.dim[V.outer])) return;
        if (C.ss && nt != ih[IH_OSTATE] / TILE_COLS) return;     // state sampling: the caches hold the sampled state's column tile only
        i1 = i2 = count_le(V.tau, n, T.beta / 2.0);
    }
    if (rare) {                                                   // whole (relabelled / shifted) sequence from the identity, no caches used
        const RareJob& J = C.rjob[chain];
        V.outer = J.new_outer; kv = 0; bv = n; i1 = 0; i2 = n;
        V.code = C.sc_code + cb;
        if (J.kind >= 4) V.tau = C.sc_tau + cb;
        store_seg = J.stored != 0;
    }
    const double* tau = V.tau; const unsigned char* code = V.code; const short* sst = V.sst; const unsigned short* slot = V.slot;
    const int outer = V.outer;
    const int s = i1, e = (rem && !rare) ? i2 + 1 : i2, lo = (rem && !rare) ? i1 + 1 : i1, hi = i2, nA = (rem || rare || meas) ? 0 : 1;
    const double tA = P.t1, tB = P.t2;
    const int codeA = P.f * 2 + P.type1, codeB = P.f * 2 + (1 - P.type1);
    const int m = ldg(&T.dim[outer]);
    const int j0 = nt * TILE_COLS, mw = m - j0 < TILE_COLS ? m - j0 : TILE_COLS;          // this job's columns
    const int nb = bv > e ? bv - e : 0, nk = kv <= s - 1 ? s - kv : 0, ns = nA + (hi - lo) + nA;
    const int total = nb + nk + ns;
    int cur = -1;                        // block of the work matrix (rows)
    int work = 0;                        // sum of d_out * d_in over the operator products of this job
    int r_opc = 0, r_sst = 0, r_slot = 0; double r_dt = 0.0;
#pragma unroll 1
    for (int j = 0; j < total; ++j) {
        if ((j & (W2B_WARP - 1)) == 0) {                       // next 32 step records, one per lane
            const int q = j + lane;
            r_opc = 0; r_sst = -1; r_slot = -1; r_dt = 0.0;
            if (q < nb) {                                       // bra step: BRA_i = O_i^T P(tau_{i+1} - tau_i) BRA_{i+1}
                const int i = bv - 1 - q;
                r_opc = (int)code[i] ^ 1; r_sst = (int)sst[i]; r_slot = (int)slot[i] | (2 << 16);
                r_dt = (i + 1 == n ? T.beta : tau[i + 1]) - tau[i];
            } else if (q < nb + nk) {                           // ket step: KET_i = O_i P(tau_i - tau_{i-1}) KET_{i-1}
                const int i = kv + (q - nb);
                r_opc = (int)code[i]; r_sst = i == 0 ? outer : (int)sst[i - 1]; r_slot = (int)slot[i] | (1 << 16);
                r_dt = tau[i] - (i == 0 ? 0.0 : tau[i - 1]);
            } else if (q < total) {                             // sub-string step: [A] old operators [lo, hi) [B]
                const int u = q - nb - nk;
                double t, tp;
                if (u < nA) { r_opc = codeA; t = tA; tp = s == 0 ? 0.0 : tau[s - 1]; }
                else if (u - nA < hi - lo) {
                    const int i = lo + u - nA; r_opc = (int)code[i]; t = tau[i]; tp = u == 0 ? (s == 0 ? 0.0 : tau[s - 1]) : (u - nA == 0 ? tA : tau[i - 1]);
                    if (store_seg) r_slot = i | (1 << 16);      // accepted tau shifts reset the slot map to the identity: operator i <-> slot i
                }
                else { r_opc = codeB; t = tB; tp = hi > lo ? tau[hi - 1] : tA; }
                r_dt = t - tp;
            }
        }
        const int src = j & (W2B_WARP - 1);
        const int opc = wbcast(r_opc, src), sin = wbcast(r_sst, src), sl = wbcast(r_slot, src);
        const double dt = wbcast(r_dt, src);
        // phase starts: load the frontier tile (or identity columns) into the work matrix
        if (j == 0 && nb > 0) job_load_bra(T, V, w, bv, j0, mw);
        else if (j == nb && nk > 0) { cur = kv == 0 ? outer : (int)sst[kv - 1]; job_load_ket(T, V, w, kv - 1, j0, mw); }
        else if (j == nb + nk && nk == 0) { cur = s == 0 ? outer : (int)sst[s - 1]; job_load_ket(T, V, w, s - 1, j0, mw); }
        const int in = sin >= 0 ? sin : cur;
        double* dst = nullptr; double* dlg = nullptr;
        if (sl >= 0) {                                          // this job's columns of the operator's cache slot are filled by the product's epilogue
            dst = ((sl >> 16) == 1 ? V.ket : V.bra) + (size_t)(sl & 0xffff) * C.mstride;
            dlg = ((sl >> 16) == 1 ? V.ketlg : V.bralg) + (size_t)(sl & 0xffff) * C.dmax + j0;
        }
        cur = tile_apply(T, in, opc, dt, w, mw, dst, dlg, j0, &work);
    }
    if (meas) return;
    if (ns == 0 && nk == 0) { cur = s == 0 ? outer : (int)sst[s - 1]; job_load_ket(T, V, w, s - 1, j0, mw); }   // nothing applied after the ket frontier: tile = KET_{s-1}
    const double tlast = ns > 0 ? (nA ? tB : tau[hi - 1]) : (s == 0 ? 0.0 : tau[s - 1]);
    double* pt = C.ptc + (size_t)chain * 2 * C.dmax + 2 * j0;
    const double* B = nullptr; const double* blg = nullptr; double tend = T.beta;         // e == n: unit bras at tau = beta
    if (e != n) {
        const int d = ldg(&T.dim[cur]);
        B = V.bra + (size_t)slot[e] * C.mstride + j0 * d; blg = V.bralg + (size_t)slot[e] * C.dmax + j0; tend = tau[e];
    }
    tile_close(T, cur, tend - tlast, w, B, blg, mw, j0, pt);
    if (lane == 0 && work) atomic_add(&C.counters[(size_t)chain * C_NCNT + C_TRFLOPS], 2ull * (unsigned long long)work * (unsigned long long)mw);   // per-chain counter: no hot address
}

// trace contributions of the outer states of an empty configuration: w_exp(-(E_j - Egs) beta)
W2B_HD void empty_parts(const Tables& T, Ctx& x, int outer) {
    const int m = ldg(&T.dim[outer]);
    const double* ev = T.evals + ldg(&T.offset[outer]);
#pragma unroll 1
    for (int j = lane_id(); j < m; j += W2B_WARP) { x.pt[2 * j] = -(ldg(&ev[j]) - T.egs) * T.beta; x.pt[2 * j + 1] = 1.0; }
    wsync();
}
// WeightedOuterStateChoice (Trace.F90:1477-1604, state sampling; no outer truncation: every superstate keeps all its states).  Picks
// the new outer (superstate, eigenstate) among the candidate superstates with probability ~ exp(-tau_new (E - Egs)) and returns
// p_prop(new -> old) / p_prop(old -> new).  Candidates: the single superstates one_old / one_new, or (flags != nullptr) every superstate
// k with flags[k] (new) / flags[nsst + k] (old).  One random number; none, and 0 returned, when nothing can be chosen.  Warp-uniform.
W2B_HDN double wosc(const Tables& T, Ctx& x, double tau_old, double tau_new, int one_old, int one_new, const int* flags, int& sst_new, int& state_new) {
    const int sst_old = x.outer, state_old = x.ostate;
    double tw = 0.0, tot = 0.0; bool any = false;
#pragma unroll 1
    for (int q = flags ? 0 : one_old; q < (flags ? T.nsst : one_old + 1); ++q) {
        if (flags && !flags[T.nsst + q]) continue;
        const double* ev = T.evals + ldg(&T.offset[q]); const int d = ldg(&T.dim[q]);
#pragma unroll 1
        for (int j = 0; j < d; ++j) tw += w_exp(-tau_old * (ldg(&ev[j]) - T.egs));
    }
    const double res = w_exp(-tau_old * (ldg(&T.evals[ldg(&T.offset[sst_old]) + state_old]) - T.egs)) / tw;
    sst_new = sst_old; state_new = state_old;
#pragma unroll 1
    for (int q = flags ? 0 : one_new; q < (flags ? T.nsst : one_new + 1); ++q) {
        if (flags && !flags[q]) continue;
        const double* ev = T.evals + ldg(&T.offset[q]); const int d = ldg(&T.dim[q]);
#pragma unroll 1
        for (int j = 0; j < d; ++j) { tot += w_exp(-tau_new * (ldg(&ev[j]) - T.egs)); any = true; }
    }
    if (!any || !(tot > 0.0)) return 0.0;
    const double rand = x.rng.next() * tot;
    double acc = 0.0;
#pragma unroll 1
    for (int q = flags ? 0 : one_new; q < (flags ? T.nsst : one_new + 1); ++q) {
        if (flags && !flags[q]) continue;
        const double* ev = T.evals + ldg(&T.offset[q]); const int d = ldg(&T.dim[q]);
#pragma unroll 1
        for (int j = 0; j < d; ++j) {
            const double w = w_exp(-tau_new * (ldg(&ev[j]) - T.egs));
            acc += w;
            if (rand <= acc && w != 0.0) { sst_new = q; state_new = j; return res * (tot / w); }
        }
    }
    return 0.0;
}
W2B_HD double outer_tau(const Tables& T, const double* tau, int n) { return n == 0 ? T.beta : T.beta - tau[n - 1] + tau[0]; }
W2B_HDN void commit_parts(const Tables& T, Ctx& x) {
    const int m = ldg(&T.dim[x.outer]);
#pragma unroll 1
    for (int j = lane_id(); j < 2 * m; j += W2B_WARP) x.pt[j] = x.mpt[j];
}

// integer superstate walk over the old operators [i0, i0+mcount) starting from `cur`; counts operators of another flavour
W2B_HDN int walk_check(const Tables& T, const Ctx& x, int i0, int mcount, int f, int cur, int& nother) {
    int oth = 0;
#pragma unroll 1
    for (int base = 0; base < mcount; base += W2B_WARP) {
        int j = base + lane_id(); bool act = j < mcount;
        int cj = act ? (int)x.code[i0 + j] : 0;
        int cnt = mcount - base < W2B_WARP ? mcount - base : W2B_WARP;
#pragma unroll 1
        for (int q = 0; q < cnt; ++q) { int cq = wbcast(cj, q); cur = cur >= 0 ? conn(T, cur, cq) : -1; }
        if (act) oth += ((cj >> 1) != f);
    }
    nother = wsum(oth);
    return cur;
}

// Every line of the k x k matrix a requested from HBM into L2 at once.  The rank-1 routines below walk the matrix in dependent
// steps of a few loads (a warp has ~4 loads in flight, one memory round trip per step: 50-100 round trips per 74 x 74 matrix);
// with the lines already on their way the same loops run at L2 latency.  Used where the matrix is about to be read whole.
W2B_HD void prefetch_matrix(const double* a, int k, int ld) {
#if W2B_DEVICE
    if (k < 16) return;
    const int nl = (k * 8 + 127) / 128 + 1;
#pragma unroll 1
    for (int c = 0; c < nl; ++c) {
#pragma unroll 1
        for (int r = lane_id(); r < k; r += W2B_WARP) {
            const char* q = (const char*)(a + (size_t)r * ld) + c * 128;
            if (q < (const char*)(a + (size_t)r * ld + k)) asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
        }
    }
#endif
}

// ------------------------------------------------------------------ determinant ratio + rank-1 updates
// S - R^T M Q (MatrixUpdate.F90:392-407); leaves Q, R, V = R^T M in the warp scratch for the update
W2B_HDN double det_ratio_add(const Tables& T, const Chains& C, Ctx& x, int f, double tau_c, double tau_a) {
    int k = x.ih[IH_KF + f];
    const double* tc = x.tc + (size_t)f * C.kcap; const double* ta = x.ta + (size_t)f * C.kcap;
#pragma unroll 1
    for (int j = lane_id(); j < k; j += W2B_WARP) {
        x.wQ[j] = hybrF(T, ta[j] - tau_c, f);              // Q_j = F(tau_a_j - tau_c,new)
        x.wR[j] = hybrF(T, tau_a - tc[j], f);              // R_i = F(tau_a,new - tau_c_i)
    }
    wsync();
    const double* M = Mbuf(C, x, f);
    const int ld = C.ld;
    double part = 0.0;
#pragma unroll 1
    for (int j = lane_id(); j < k; j += W2B_WARP) {
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        int i = 0;
#pragma unroll 1
        for (; i + 3 < k; i += 4) {
            a0 += x.wR[i] * M[(size_t)i * ld + j];
            a1 += x.wR[i + 1] * M[(size_t)(i + 1) * ld + j];
            a2 += x.wR[i + 2] * M[(size_t)(i + 2) * ld + j];
            a3 += x.wR[i + 3] * M[(size_t)(i + 3) * ld + j];
        }
#pragma unroll 1
        for (; i < k; ++i) a0 += x.wR[i] * M[(size_t)i * ld + j];
        double v = (a0 + a1) + (a2 + a3);
        x.wV[j] = v;
        part += v * x.wQ[j];
    }
    double S = hybrF(T, tau_a - tau_c, f) - wsum(part);
    wsync();
    return S;
}

// get_MatAdd MatrixUpdate.F90:410-442, fused with the row/column shift, written into the spare buffer.
// Two passes (profiles/r2q_*: the row-at-a-time form -- load a row, reduce, store it -- spent 70-80 % of its samples waiting for
// one row's loads at a time): u = M Q for four rows per iteration (twelve loads in flight per lane), then the rank-1 update
// streams two rows at a time with nothing but independent loads and stores.  Same operations per element as before.
W2B_HDN void mat_add(const Chains& C, Ctx& x, int f, int nf, int p1, int p2, double S) {
    int k = x.ih[IH_KF + f];
    const double* src = Mbuf(C, x, f); double* dst = Mbuf(C, x, nf);
    const int ld = C.ld; const double inv = 1.0 / S;
    double* wU = x.wR;                                                 // R is spent once V = R^T M exists
#pragma unroll 1
    for (int i0 = 0; i0 < k; i0 += 4) {
        const double* r0 = src + (size_t)i0 * ld;
        const double* r1 = i0 + 1 < k ? r0 + ld : r0; const double* r2 = i0 + 2 < k ? r0 + 2 * (size_t)ld : r0; const double* r3 = i0 + 3 < k ? r0 + 3 * (size_t)ld : r0;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 1
        for (int j = lane_id(); j < k; j += W2B_WARP) { const double q = x.wQ[j]; s0 += r0[j] * q; s1 += r1[j] * q; s2 += r2[j] * q; s3 += r3[j] * q; }
        s0 = wsum(s0); s1 = wsum(s1); s2 = wsum(s2); s3 = wsum(s3);
        if (lane_id() == 0) { wU[i0] = s0; if (i0 + 1 < k) wU[i0 + 1] = s1; if (i0 + 2 < k) wU[i0 + 2] = s2; if (i0 + 3 < k) wU[i0 + 3] = s3; }
    }
    wsync();
#pragma unroll 1
    for (int jn = lane_id(); jn <= k; jn += W2B_WARP)
        dst[(size_t)p1 * ld + jn] = (jn == p2) ? inv : -x.wV[jn - (jn > p2)] * inv;
#pragma unroll 1
    for (int i = 0; i < k; i += 2) {
        const bool two = i + 1 < k;
        const double* ra = src + (size_t)i * ld; const double* rb = two ? ra + ld : ra;
        double* da = dst + (size_t)(i + (i >= p1)) * ld; double* db = dst + (size_t)(i + 1 + (i + 1 >= p1)) * ld;
        const double ua = wU[i], ub = two ? wU[i + 1] : 0.0;
#pragma unroll 1
        for (int jn = lane_id(); jn <= k; jn += W2B_WARP) {
            if (jn == p2) { da[jn] = -ua * inv; if (two) db[jn] = -ub * inv; }
            else {
                const int j = jn - (jn > p2);
                const double v = x.wV[j], a = ra[j], b = rb[j];
                da[jn] = a + ua * v * inv;
                if (two) db[jn] = b + ub * v * inv;
            }
        }
    }
    wsync();
    if (lane_id() == 0) { int t = x.ih[IH_MSLOT + f]; x.ih[IH_MSLOT + f] = x.ih[IH_MSLOT + nf]; x.ih[IH_MSLOT + nf] = t; x.ih[IH_KF + f] = k + 1; }
    wsync();
}

// get_MatRem MatrixUpdate.F90:455-482 (two rows per iteration: independent loads in flight)
W2B_HDN void mat_rem(const Chains& C, Ctx& x, int f, int nf, int p1, int p2) {
    int k = x.ih[IH_KF + f];
    const double* src = Mbuf(C, x, f); double* dst = Mbuf(C, x, nf);
    const int ld = C.ld;
#pragma unroll 1
    for (int j = lane_id(); j < k; j += W2B_WARP) { x.wV[j] = src[(size_t)p1 * ld + j]; x.wQ[j] = src[(size_t)j * ld + p2]; }
    wsync();
    const double ipiv = 1.0 / src[(size_t)p1 * ld + p2];
#pragma unroll 1
    for (int i = 0; i < k - 1; i += 2) {
        const bool two = i + 1 < k - 1;
        const int ia = i + (i >= p1), ib = i + 1 + (i + 1 >= p1);
        const double ca = x.wQ[ia], cb = two ? x.wQ[ib] : 0.0;
        const double* ra = src + (size_t)ia * ld; const double* rb = two ? src + (size_t)ib * ld : ra;
        double* da = dst + (size_t)i * ld; double* db = da + ld;
#pragma unroll 1
        for (int j = lane_id(); j < k - 1; j += W2B_WARP) {
            const int jo = j + (j >= p2);
            const double v = x.wV[jo], a = ra[jo], b = rb[jo];
            da[j] = a - ca * v * ipiv;
            if (two) db[j] = b - cb * v * ipiv;
        }
    }
    wsync();
    if (lane_id() == 0) { int t = x.ih[IH_MSLOT + f]; x.ih[IH_MSLOT + f] = x.ih[IH_MSLOT + nf]; x.ih[IH_MSLOT + nf] = t; x.ih[IH_KF + f] = k - 1; }
    wsync();
}

// ------------------------------------------------------------------ logging helpers
// (compiling the replay log out was measured: -5 % on k_propose, nothing on k_decide -- it stays in)
W2B_HD void log_begin(Ctx& x, int type) {
    if (x.log && lane_id() == 0) { for (int i = 0; i < LOGW; ++i) x.log[i] = 0.0; x.log[0] = type; x.log[4] = -1; }
}
W2B_HD void log_set(Ctx& x, int i, double v) { if (x.log && lane_id() == 0) x.log[i] = v; }
W2B_HD void log_end(Ctx& x, int status) {
    if (x.log && lane_id() == 0) { x.log[1] = status; x.log[12] = x.n; x.log[13] = x.outer; x.log[15] = x.cfgsign; }
}
W2B_HD void bump(Ctx& x, int c) { if (lane_id() == 0) x.cnt[c] += 1ull; }
// algorithmic bytes of the determinant path (definition in DESIGN.md "Roofline accounting")
W2B_HD void bump_bytes(Ctx& x, long long b) { if (lane_id() == 0) x.cnt[C_ALGBYTES] += (unsigned long long)b; }

// ------------------------------------------------------------------ pair insertion / removal, split into two phases
// The CTA-level scheduler (cta_sweep below) runs all chains of a CTA through the SAME phase at the same time so that the
// warps of an SM share instruction-cache lines:  propose (RNG draws, positions, quantum-number walk, scalar trace ratio)
// -> decide (determinant ratio, blocked trace, Metropolis draw, rank-1 update).  The proposal travels in a Prop record.

// StepAdd part 1 (CTQMC.F90:3353-3469): pair_OperAdd, process_OperAdd, check_sst_sequence.  Returns false on an early exit.
// OUTER = change_outer (PercentageOuterMove, CTQMC.F90:3422-3449): the superstates BETWEEN the two new operators are kept, the part of
// the sequence from the later operator over tau = beta to the earlier one is re-walked, and the superstate it passes beta with
// becomes the outer superstate (check_sst_sequence, Trace.F90:1278-1295); it is left in P.begin_sst.  These moves are served like
// global moves (whole-sequence trace from scratch: CalcNew, Trace.F90:5381), never by the hot propose kernel.
template <bool OUTER>
W2B_HDN bool add_propose(const Tables& T, const Run& R, const Chains& C, Ctx& x, Prop& P) {
    log_begin(x, OUTER ? 5 : 1);
    bump(x, C_TRYADD);
    const double W = x.wmax - x.wmin;
    const double tau_a = add_rn(x.wmin, mul_rn(W, x.rng.next()));                    // pair_OperAdd Trace.F90:3672-3695
    const int b = x.rng.randint(1, T.nb) - 1;
    const int s = x.rng.randint(1, 2) - 1;
    const double tmin = fmax(tau_a - R.taudiff_max, x.wmin);
    const double w2 = fmin(tau_a + R.taudiff_max, x.wmax) - tmin;
    const double tau_c = add_rn(tmin, mul_rn(w2, x.rng.next()));
    const int f = b * 2 + s;
    const bool a_first = tau_a < tau_c;
    const double t1 = a_first ? tau_a : tau_c, t2 = a_first ? tau_c : tau_a;
    const int code1 = f * 2 + (a_first ? 1 : 0), code2 = f * 2 + (a_first ? 0 : 1);
    log_set(x, 4, f); log_set(x, 2, t1); log_set(x, 3, t2);
    bool eq12; int i1, i2;
    count_lt2(x.tau, x.n, t1, t2, i1, i2, eq12);
    if (eq12 || tau_a == tau_c) { bump(x, C_QNADD); log_end(x, 2); return false; }     // process_OperAdd: equal times
    const int k = x.ih[IH_KF + f];
    const double* tc = x.tc + (size_t)f * C.kcap; const double* ta = x.ta + (size_t)f * C.kcap;
    const int p1 = count_lt(tc, k, tau_c), p2 = count_lt(ta, k, tau_a);              // FPos-1 (creator row, annihilator column)
    log_set(x, 5, p1 + 1); log_set(x, 6, p2 + 1);
    bump_bytes(x, 64ll * k);
    // quantum-number check of the changed sub-string (+ scalar trace ratio when every block is 1x1)
    int begin_sst = i1 == 0 ? x.outer : (int)x.sst[i1 - 1];
    double dlog = 0.0, dsign = 1.0;
    int cur = conn(T, begin_sst, code1), old_after = begin_sst, nother = 0;
    double tl = t1;
    int new_end = -1;
    if (OUTER) {
        const int s_b = i2 == 0 ? x.outer : (int)x.sst[i2 - 1];     // superstate the later operator acts on; begin_sst = the one the earlier has to produce
        int dummy;
        cur = conn(T, s_b, code2);
        if (cur >= 0) cur = walk_check(T, x, i2, x.n - i2, f, cur, dummy);
        const int new_outer = cur;
        if (cur >= 0) cur = walk_check(T, x, 0, i1, f, cur, dummy);
        if (cur >= 0) cur = conn(T, cur, code1);
        walk_check(T, x, i1, i2 - i1, f, 0, nother);                // operators of other flavours between the two (permutation sign)
        new_end = cur; old_after = begin_sst; begin_sst = new_outer;
    } else if (T.dmax == 1) {
        if (cur >= 0) {
            double e = ldg(&T.elem[begin_sst * T.nops + code1]); dlog += ldg(&T.lelem[begin_sst * T.nops + code1]); if (e < 0) dsign = -dsign;
            walk_delta(T, x, i1, i2 - i1, f, cur, old_after, tl, dlog, dsign, nother);
        }
        if (cur >= 0) {
            dlog += -(ldg(&T.e0[cur]) - ldg(&T.e0[old_after])) * (t2 - tl);
            double e = ldg(&T.elem[cur * T.nops + code2]);
            new_end = conn(T, cur, code2);
            if (new_end >= 0) { dlog += ldg(&T.lelem[cur * T.nops + code2]); if (e < 0) dsign = -dsign; }
        }
    } else {
        if (cur >= 0) cur = walk_check(T, x, i1, i2 - i1, f, cur, nother);
        if (i2 > i1) old_after = (int)x.sst[i2 - 1];
        if (cur >= 0) new_end = conn(T, cur, code2);
    }
    if (new_end < 0 || new_end != old_after) { bump(x, C_QNADD); log_end(x, 2); return false; }
    if (lane_id() == 0) {
        P.t1 = t1; P.t2 = t2; P.tau_a = tau_a; P.tau_c = tau_c; P.factor = mul_rn(W, w2); P.dlog = dlog; P.dsign = dsign;
        P.f = f; P.i1 = i1; P.i2 = i2; P.p1 = p1; P.p2 = p2; P.begin_sst = begin_sst; P.nother = nother; P.type1 = a_first ? 1 : 0;
    }
    wsync();
    return true;
}

// StepAdd part 2 (CTQMC.F90:3470-3612): determinant ratio, local trace, Metropolis, get_MatAdd + list insertion
// OUTER: the proposed sequence (times, codes, superstates from the new outer superstate P.begin_sst) is staged in the scratch arrays
// (outer_stage) and its local trace was evaluated from scratch: parts in x.mpt (blocked) / ratio in P.dlog, P.dsign (1x1 blocks).
template <bool OUTER>
W2B_HDN void add_decide(const Tables& T, const Run& R, const Chains& C, const Accum& A, Ctx& x, const Prop& P) {
    const int f = P.f, i1 = P.i1, i2 = P.i2, p1 = P.p1, p2 = P.p2, begin_sst = P.begin_sst, nother = P.nother;
    const double t1 = P.t1, t2 = P.t2, tau_a = P.tau_a, tau_c = P.tau_c;
    const int code1 = f * 2 + P.type1, code2 = f * 2 + (1 - P.type1);
    const bool blocked = T.dmax > 1;
    const int k = x.ih[IH_KF + f];
    const double* tc = x.tc + (size_t)f * C.kcap; const double* ta = x.ta + (size_t)f * C.kcap;
    double dlog = P.dlog, dsign = P.dsign;
    prefetch_matrix(Mbuf(C, x, f), k, C.ld);
    const double S = det_ratio_add(T, C, x, f, tau_c, tau_a);
    bump_bytes(x, 8ll * k * k);
    double tn_log = 0.0, tn_sign = 1.0;
    if (blocked) {                                                     // get_trace_EB of the proposed configuration
        // the trace kernel has left the per-outer-state parts of the proposed configuration in C.ptc and extended the cache frontiers
        parts_sum(x.mpt, ldg(&T.dim[OUTER ? begin_sst : x.outer]), tn_log, tn_sign, OUTER ? (R.statesampling ? P.new_state : -1) : x.only);
        if (!OUTER) { if (x.kv < i1) x.kv = i1; if (x.bv > i2) x.bv = i2; }
        dlog = tn_log - x.tr_log; dsign = tn_sign * x.tr_sign;
    }
    const double sig = ((p1 - p2) & 1) ? -1.0 : 1.0;
    const double dr_log = w_log(fabs(S)), dr_sign = sgn1(S) * sig;
    const double rand = x.rng.next();
    // NPairs after insertion (CTQMC.F90:3510-3515)
    if (x.npairs < 0) x.npairs = count_pairs_full(T, R, C, x);
    const int np_new = x.npairs + pairs_with(R, x, ta, k, tau_c) + pairs_with(R, x, tc, k, tau_a) + (pairable(R, x, tau_a, tau_c) ? 1 : 0);
    // state sampling, outer moves: P.factor carries the WeightedOuterStateChoice factor (CTQMC.F90:3471-3480, 3527); 0 = no outer state
    // could be chosen: rejected whatever the trace
    const bool dead = OUTER && R.statesampling && P.factor == 0.0;
    if (dead) { dlog = 0.0; dsign = 0.0; }
    const double prob = dead ? 0.0 : fabs(P.factor * (double)(2 * T.nb) / (double)np_new * w_exp(dlog + dr_log));
    log_set(x, 7, dr_log); log_set(x, 8, dr_sign); log_set(x, 9, dlog); log_set(x, 10, dsign); log_set(x, 11, prob);
    log_set(x, 14, np_new); log_set(x, 17, rand);
    if (!(rand < prob)) { log_end(x, 0); return; }
    if (k + 1 > C.kcap || x.n + 2 > C.ops_cap) { x.err |= ERR_CAPACITY; log_end(x, 0); return; }
    // ---- accept
    bump_bytes(x, 8ll * (k * k + (k + 1) * (k + 1)));
    mat_add(C, x, f, T.nf, p1, p2, S);
    shift_up(x.tc + (size_t)f * C.kcap, p1, k, 1); shift_up(x.ta + (size_t)f * C.kcap, p2, k, 1);
    if (lane_id() == 0) { x.tc[(size_t)f * C.kcap + p1] = tau_c; x.ta[(size_t)f * C.kcap + p2] = tau_a; }
    if (OUTER) {                                                       // the staged sequence replaces the old one; every cached product is void
#pragma unroll 1
        for (int i = lane_id(); i < x.n + 2; i += W2B_WARP) { x.tau[i] = x.sc_tau[i]; x.code[i] = x.sc_code[i]; x.sst[i] = x.sc_sst[i]; }
        wsync();
        x.n += 2; x.outer = begin_sst;
        if (R.statesampling) { x.ostate = P.new_state; x.only = x.ostate; }
        if (blocked) { commit_parts(T, x); slots_reset(C, x); }
    } else {
    shift_up(x.tau, i2, x.n, 2); shift_up(x.code, i2, x.n, 2); shift_up(x.sst, i2, x.n, 2);
    shift_up(x.tau, i1, i2, 1); shift_up(x.code, i1, i2, 1); shift_up(x.sst, i1, i2, 1);
    if (lane_id() == 0) { x.tau[i1] = t1; x.code[i1] = (unsigned char)code1; x.tau[i2 + 1] = t2; x.code[i2 + 1] = (unsigned char)code2; }
    if (blocked) {                                                     // two fresh cache slots; kets valid below i1, bras from i2+2 on
        shift_up(x.slot, i2, x.n, 2); shift_up(x.slot, i1, i2, 1);
        if (lane_id() == 0) { x.slot[i1] = x.freel[x.nfree - 1]; x.slot[i2 + 1] = x.freel[x.nfree - 2]; }
        x.nfree -= 2; x.kv = i1; x.bv = i2 + 2;
        commit_parts(T, x);
    }
    wsync();
    x.n += 2;
    walk_store(T, x.code, i1, i2 - i1 + 2, begin_sst, x.sst, i1);
    }
    x.det_log += dr_log; x.det_sign *= dr_sign;
    if (blocked) { x.tr_log = tn_log; x.tr_sign = tn_sign; } else { x.tr_log += dlog; x.tr_sign *= dsign; }
    x.perm ^= (nother & 1);
    x.palt = (x.palt & ~(1 << f)) | (palt_flavour(C, x, f) << f);
    x.npairs = np_new;
    update_cfgsign(T, x);
    bump(x, C_ACCADD);
    if (lane_id() == 0) {
        int bin = (int)floor((t2 - t1) / R.accpair_max * (double)R.naccpair) + 1; if (bin > R.naccpair) bin = R.naccpair;   // CTQMC.F90:3579-3582
        atomic_add(&A.accpair[(size_t)f * R.naccpair + bin - 1], 1ull);
        if (A.pairlist_cap > 0) { const int i = atomic_add_int(A.pairlist_cnt, 1); if (i < A.pairlist_cap) A.pairlist[i] = t2 - t1; }
    }
    log_end(x, 1);
}

// StepRem part 1 (CTQMC.F90:3892-3977): gen_OperRemove (pair choice in the reference's enumeration order), quantum numbers
// OUTER = change_outer (CTQMC.F90:3934-3961), see add_propose: the new outer superstate is left in P.begin_sst.
template <bool OUTER>
W2B_HDN bool rem_propose(const Tables& T, const Run& R, const Chains& C, Ctx& x, Prop& P) {
    log_begin(x, OUTER ? 6 : 2);
    bump(x, C_TRYREM);
    if (x.npairs < 0) x.npairs = count_pairs_full(T, R, C, x);        // gen_OperRemove Trace.F90:4868-4877
    log_set(x, 14, x.npairs);
    if (x.npairs <= 0) { bump(x, C_QNREM); log_end(x, 2); return false; }
    const int r = x.rng.randint(1, x.npairs);
    const int ia = count_lt(x.tau, x.n, x.wmin), ib = count_le(x.tau, x.n, x.wmax);
    // enumerate (El, mEl) in the reference's order: El over in-window operators in time order, mEl over later
    // same-flavour operators of opposite type with tau <= wmax and tau <= tau_El + taudiff_max
    int i1 = -1, f = 0, type1 = 0, midx = 0; double t1 = 0.0;
    int cum = 0;
#pragma unroll 1
    for (int base = ia; base < ib; base += W2B_WARP) {
        int i = base + lane_id(); bool act = i < ib;
        int c = 0, lb = 0, fi = 0, ty = 0; double ti = 0.0;
        if (act) {
            int cd = x.code[i]; fi = cd >> 1; ty = cd & 1; ti = x.tau[i];
            int k = x.ih[IH_KF + fi];
            const double* opp = (ty == 0 ? x.ta : x.tc) + (size_t)fi * C.kcap;
            {                                                     // lb = number of opposite-type times <= ti (sorted): binary search
                int lo = 0, hi = k;
#pragma unroll 1
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (opp[mid] <= ti) lo = mid + 1; else hi = mid; }
                lb = lo;
            }
#pragma unroll 1
            for (int a = lb; a < k; ++a) {
                const double v = opp[a];
                if (v <= x.wmax && !(v > ti + R.taudiff_max)) ++c;
                else break;
            }
        }
        int incl = wscan_incl(c);
        int total = wbcast(incl, W2B_WARP - 1);
        if (cum + total >= r) {
            unsigned hit = wballot(act && (cum + incl >= r));
            int src = 0;
#if W2B_DEVICE
            src = __ffs(hit) - 1;
#endif
            int rr = r - (cum + incl - c);
            i1 = base + src; f = wbcast(fi, src); type1 = wbcast(ty, src); t1 = wbcast(ti, src); midx = wbcast(lb + rr - 1, src);
            break;
        }
        cum += total;
    }
    if (i1 < 0) { x.err |= ERR_PAIR_LOGIC; log_end(x, 2); return false; }
    const int k = x.ih[IH_KF + f];
    const double* tc = x.tc + (size_t)f * C.kcap; const double* ta = x.ta + (size_t)f * C.kcap;
    const double t2 = (type1 == 0 ? ta : tc)[midx];
    bump_bytes(x, 64ll * k);
    const int pE = count_lt(type1 == 0 ? tc : ta, k, t1);
    const int p1 = type1 == 0 ? pE : midx, p2 = type1 == 0 ? midx : pE;
    const double tau_an = type1 == 1 ? t1 : t2;
    const int i2 = count_lt(x.tau, x.n, t2);
    log_set(x, 2, t1); log_set(x, 3, t2); log_set(x, 4, f); log_set(x, 5, p1 + 1); log_set(x, 6, p2 + 1);
    const double factor = (double)x.npairs / ((x.wmax - x.wmin) * (fmin(tau_an + R.taudiff_max, x.wmax) - fmax(tau_an - R.taudiff_max, x.wmin)));
    // quantum numbers (+ scalar trace ratio) of the configuration without the pair
    const int code1 = f * 2 + type1, code2 = f * 2 + (1 - type1);
    int begin_sst = i1 == 0 ? x.outer : (int)x.sst[i1 - 1];
    double dlog = 0.0, dsign = 1.0;
    int cur = begin_sst, old_after = (int)x.sst[i1], nother = 0;
    double tl = t1;
    const int end_sst = (int)x.sst[i2];
    if (OUTER) {
        int dummy;
        cur = walk_check(T, x, i2 + 1, x.n - i2 - 1, f, (int)x.sst[i2 - 1], dummy);      // from the superstate before the later operator over beta ...
        const int new_outer = cur;
        if (cur >= 0) cur = walk_check(T, x, 0, i1, f, cur, dummy);                       // ... to the superstate after the earlier one
        if (cur < 0 || cur != old_after) { bump(x, C_QNREM); log_end(x, 2); return false; }
        walk_check(T, x, i1 + 1, i2 - i1 - 1, f, 0, nother);
        begin_sst = new_outer;
    } else if (T.dmax == 1) {
        { double e = ldg(&T.elem[begin_sst * T.nops + code1]); dlog -= ldg(&T.lelem[begin_sst * T.nops + code1]); if (e < 0) dsign = -dsign; }
        walk_delta(T, x, i1 + 1, i2 - i1 - 1, f, cur, old_after, tl, dlog, dsign, nother);
        if (cur < 0 || cur != end_sst) { bump(x, C_QNREM); log_end(x, 2); return false; }
        dlog += -(ldg(&T.e0[cur]) - ldg(&T.e0[old_after])) * (t2 - tl);
        { double e = ldg(&T.elem[old_after * T.nops + code2]); dlog -= ldg(&T.lelem[old_after * T.nops + code2]); if (e < 0) dsign = -dsign; }
    } else {
        cur = walk_check(T, x, i1 + 1, i2 - i1 - 1, f, cur, nother);
        if (cur < 0 || cur != end_sst) { bump(x, C_QNREM); log_end(x, 2); return false; }
    }
    if (lane_id() == 0) {
        P.t1 = t1; P.t2 = t2; P.tau_a = 0.0; P.tau_c = 0.0; P.factor = factor; P.dlog = dlog; P.dsign = dsign;
        P.f = f; P.i1 = i1; P.i2 = i2; P.p1 = p1; P.p2 = p2; P.begin_sst = begin_sst; P.nother = nother; P.type1 = type1;
    }
    wsync();
    return true;
}

// StepRem part 2 (CTQMC.F90:3978-4150): ratio = M(p1,p2), local trace, Metropolis, get_MatRem + list removal
template <bool OUTER>
W2B_HDN void rem_decide(const Tables& T, const Run& R, const Chains& C, const Accum& A, Ctx& x, const Prop& P) {
    const int f = P.f, i1 = P.i1, i2 = P.i2, p1 = P.p1, p2 = P.p2, begin_sst = P.begin_sst, nother = P.nother, type1 = P.type1;
    const double t1 = P.t1, t2 = P.t2;
    const bool blocked = T.dmax > 1;
    const int k = x.ih[IH_KF + f];
    const double* tc = x.tc + (size_t)f * C.kcap; const double* ta = x.ta + (size_t)f * C.kcap;
    double dlog = P.dlog, dsign = P.dsign, tn_log = 0.0, tn_sign = 1.0;
    prefetch_matrix(Mbuf(C, x, f), k, C.ld);                            // read whole if the removal is accepted (1 in 5 at C5; HBM has the headroom)
    if (blocked) {
        parts_sum(x.mpt, ldg(&T.dim[OUTER ? begin_sst : x.outer]), tn_log, tn_sign, OUTER ? (R.statesampling ? P.new_state : -1) : x.only);
        if (!OUTER) { if (x.kv < i1) x.kv = i1; if (x.bv > i2 + 1) x.bv = i2 + 1; }
        dlog = tn_log - x.tr_log; dsign = tn_sign * x.tr_sign;
    }
    // determinant ratio = M(p1,p2)  (MatrixUpdate.F90:445-452)
    const double S = Mbuf(C, x, f)[(size_t)p1 * C.ld + p2];
    const double sig = ((p1 - p2) & 1) ? -1.0 : 1.0;
    const double dr_log = w_log(fabs(S)), dr_sign = sgn1(S) * sig;
    const bool dead = OUTER && R.statesampling && P.factor == 0.0;              // CTQMC.F90:3979-3996, see add_decide
    if (dead) { dlog = 0.0; dsign = 0.0; }
    const double prob = dead ? 0.0 : fabs(P.factor / (double)(2 * T.nb) * w_exp(dlog + dr_log));
    const double rand = x.rng.next();
    log_set(x, 7, dr_log); log_set(x, 8, dr_sign); log_set(x, 9, dlog); log_set(x, 10, dsign); log_set(x, 11, prob); log_set(x, 17, rand);
    if (!(rand < prob)) { log_end(x, 0); return; }
    if (fabs(S) < 1e-12) x.err |= ERR_DETRAT_SMALL;                    // CTQMC.F90:4062 `stop`
    // ---- accept
    const int np_new = x.npairs - (pairs_with(R, x, type1 == 0 ? ta : tc, k, t1) + pairs_with(R, x, type1 == 0 ? tc : ta, k, t2) - (pairable(R, x, t1, t2) ? 1 : 0));
    bump_bytes(x, 8ll * (k * k + (k - 1) * (k - 1)));
    mat_rem(C, x, f, T.nf, p1, p2);
    shift_down(x.tc + (size_t)f * C.kcap, p1, k, 1); shift_down(x.ta + (size_t)f * C.kcap, p2, k, 1);
    if (OUTER) {
#pragma unroll 1
        for (int i = lane_id(); i < x.n - 2; i += W2B_WARP) { x.tau[i] = x.sc_tau[i]; x.code[i] = x.sc_code[i]; x.sst[i] = x.sc_sst[i]; }
        wsync();
        x.n -= 2; x.outer = begin_sst;
        if (R.statesampling) { x.ostate = P.new_state; x.only = x.ostate; }
        if (blocked) { commit_parts(T, x); slots_reset(C, x); }
    } else {
    if (blocked) {                                                     // give the two cache slots back; kets valid below i1, bras from i2-1 on
        if (lane_id() == 0) { x.freel[x.nfree] = x.slot[i1]; x.freel[x.nfree + 1] = x.slot[i2]; }
        wsync();
        shift_down(x.slot, i1, i2, 1); shift_down(x.slot, i2 - 1, x.n, 2);
        x.nfree += 2; x.kv = i1; x.bv = i2 - 1;
        commit_parts(T, x);
    }
    shift_down(x.tau, i1, i2, 1); shift_down(x.code, i1, i2, 1); shift_down(x.sst, i1, i2, 1);
    shift_down(x.tau, i2 - 1, x.n, 2); shift_down(x.code, i2 - 1, x.n, 2); shift_down(x.sst, i2 - 1, x.n, 2);
    x.n -= 2;
    walk_store(T, x.code, i1, i2 - i1 - 1, begin_sst, x.sst, i1);
    }
    x.det_log += dr_log; x.det_sign *= dr_sign;
    if (blocked) { x.tr_log = tn_log; x.tr_sign = tn_sign; } else { x.tr_log += dlog; x.tr_sign *= dsign; }
    x.perm ^= (nother & 1);
    x.palt = (x.palt & ~(1 << f)) | (palt_flavour(C, x, f) << f);
    x.npairs = np_new;
    update_cfgsign(T, x);
    bump(x, C_ACCREM);
    if (lane_id() == 0) {
        int bin = (int)floor((t2 - t1) / R.accpair_max * (double)R.naccpair) + 1; if (bin > R.naccpair) bin = R.naccpair;   // CTQMC.F90:3579-3582
        atomic_add(&A.accpair[(size_t)f * R.naccpair + bin - 1], 1ull);
        if (A.pairlist_cap > 0) { const int i = atomic_add_int(A.pairlist_cnt, 1); if (i < A.pairlist_cap) A.pairlist[i] = t2 - t1; }
    }
    log_end(x, 1);
}

// ------------------------------------------------------------------ full (global) evaluations
// scalar local trace of a whole configuration given codes/taus/ssts (get_Trace_EB_onlyscalars Trace.F90:5556-5727)
W2B_HDN void trace_full_scalar(const Tables& T, const double* tau, const unsigned char* code, const short* sst, int n, int outer,
                              double& lg, double& sg) {
    double acc = 0.0; int neg = 0;
#pragma unroll 1
    for (int i = lane_id(); i < n; i += W2B_WARP) {
        int before = i == 0 ? outer : (int)sst[i - 1];
        double tp = i == 0 ? 0.0 : tau[i - 1];
        acc += -(ldg(&T.e0[before]) - T.egs) * (tau[i] - tp);
        double e = ldg(&T.elem[before * T.nops + code[i]]);
        acc += ldg(&T.lelem[before * T.nops + code[i]]); neg += (e < 0.0);
    }
    double tlast = n > 0 ? tau[n - 1] : 0.0;
    lg = wsum(acc) - (ldg(&T.e0[outer]) - T.egs) * (T.beta - tlast);
    sg = (wsum(neg) & 1) ? -1.0 : 1.0;
}

// In-place Gauss-Jordan inversion with partial pivoting of the k x k row-major matrix a (leading dimension ld);
// returns log|det| and sign through the references; singular -> sign 0.  Uses wI[0..k) as pivot bookkeeping.
// (stands in for DGETRF/DGETRI of get_MatLogDetFull, MatrixUpdate.F90:41-99)
// colv / rowv: two warp-scratch vectors of k doubles.  Per pivot the elimination factors (column c) and the scaled pivot row
// are staged there once; the k row updates that follow are then independent of each other (a lane only touches its own
// columns), so their loads and stores pipeline instead of paying one global-memory round trip and two warp barriers per row
// (a 100 x 100 inverse took 5-10 ms on one warp that way and dominated the beta = 200 workload).
W2B_HDN void invert_inplace(double* a, int k, int ld, int* wI, double* colv, double* rowv, bool want_inverse, double& lg, double& sg) {
    lg = 0.0; sg = 1.0;
    const int lane = lane_id();
#pragma unroll 1
    for (int c = 0; c < k; ++c) {
        // pivot search in column c among rows that have not been used as pivot rows (rows >= c after swapping)
        double best = 0.0; int bi = c;
#pragma unroll 1
        for (int r = c + lane; r < k; r += W2B_WARP) { double v = fabs(a[(size_t)r * ld + c]); if (v > best) { best = v; bi = r; } }
        double bv; int p = wargmax_abs(best, bi, bv);
        if (!(bv > 0.0)) { sg = 0.0; lg = -1.7976931348623157e308; return; }
        if (lane == 0) wI[c] = p;
        if (p != c) {
#pragma unroll 1
            for (int j = lane; j < k; j += W2B_WARP) { double t = a[(size_t)c * ld + j]; a[(size_t)c * ld + j] = a[(size_t)p * ld + j]; a[(size_t)p * ld + j] = t; }
            sg = -sg;
        }
        wsync();
        double piv = a[(size_t)c * ld + c];
        lg += w_log(fabs(piv)); if (piv < 0) sg = -sg;
        double ipiv = 1.0 / piv;
        wsync();
        if (want_inverse) {
            // Gauss-Jordan step: row c scaled, eliminated from every other row; column c becomes the inverse column
#pragma unroll 1
            for (int j = lane; j < k; j += W2B_WARP) { const double v = (j == c) ? ipiv : a[(size_t)c * ld + j] * ipiv; a[(size_t)c * ld + j] = v; rowv[j] = v; }
#pragma unroll 1
            for (int r = lane; r < k; r += W2B_WARP) colv[r] = a[(size_t)r * ld + c];
            wsync();
#pragma unroll 1
            for (int r = 0; r < k; ++r) {
                if (r == c) continue;
                const double fct = colv[r];
#pragma unroll 1
                for (int j = lane; j < k; j += W2B_WARP)
                    a[(size_t)r * ld + j] = (j == c) ? -fct * ipiv : a[(size_t)r * ld + j] - fct * rowv[j];
            }
            wsync();
        } else {
#pragma unroll 1
            for (int j = c + 1 + lane; j < k; j += W2B_WARP) rowv[j] = a[(size_t)c * ld + j];
#pragma unroll 1
            for (int r = c + 1 + lane; r < k; r += W2B_WARP) colv[r] = a[(size_t)r * ld + c] * ipiv;
            wsync();
#pragma unroll 1
            for (int r = c + 1; r < k; ++r) {
                const double fct = colv[r];
#pragma unroll 1
                for (int j = c + 1 + lane; j < k; j += W2B_WARP) a[(size_t)r * ld + j] -= fct * rowv[j];
            }
            wsync();
        }
    }
    if (want_inverse) {
        // undo the row swaps as column swaps in reverse order: inv(P A) = inv(A) P^T
#pragma unroll 1
        for (int c = k - 1; c >= 0; --c) {
            int p = wI[c];
            if (p != c) for (int r = lane; r < k; r += W2B_WARP) { double t = a[(size_t)r * ld + c]; a[(size_t)r * ld + c] = a[(size_t)r * ld + p]; a[(size_t)r * ld + p] = t; }
            wsync();
        }
    }
}

// Blocked Gauss-Jordan with partial pivoting on the FP64 tensor pipe: the same job as invert_inplace (in-place inverse, or only
// log|det| / sign, of the k x k row-major matrix a) for the large matrices of global moves at large beta, where the k^3 work of
// DGETRF / DGETRI (MatrixUpdate.F90:41-99) is most of the arithmetic of the whole run (C5: every global move is accepted, so each
// needs 2 determinants and 2 inverses of 74 x 74 matrices -- 70 % of all flops; the unblocked warp version re-read and re-wrote
// the matrix once per pivot, a warp GROUP eliminating it in shared memory spent 47 % of the sweep kernel's instructions there and
// served two matrices per SM at a time).  Here ONE warp owns a matrix, in global memory (L1 / L2), no CTA barriers:
//   panel  4 pivots at a time: unblocked Gauss-Jordan steps restricted to the panel's 4 columns (pivot search, row swap, lanes over
//          rows).  Afterwards the panel holds P = [ A_BB^-1 ; -A_OB A_BB^-1 ] (B = the 4 pivot rows, O = the other rows).
//   update every other column at once: rows O  A_OO += P_OB R,  rows B  A_BO = P_BB R  with R = the 4 pivot rows before the update --
//          one DMMA.8x8x4 per 8 x 8 tile (K = 4 = the panel width), 16 tiles' worth of data touched once per panel instead of once
//          per pivot.
// Determinant only (want_inverse = false): panel and update are restricted to the trailing rows / columns (an LU in disguise).
// In exact arithmetic the pivots are those of the unblocked elimination; the determinant is accumulated as a product with the
// exponent split off (one logarithm per matrix instead of one per pivot).
constexpr int LU_BLOCKED_MIN = 24;     // orders from which rebuild_inverses uses the blocked routine
W2B_HDN void invert_blocked(double* a, int k, int ld, int* wI, bool want_inverse, double& lg, double& sg) {
    // (Measured and rejected: the panel held in registers, lane l owning rows l, l + 32, l + 64, with the A fragments shuffled out of
    //  them -- no memory traffic inside the panel, yet 60 % slower per global move: the selects and shuffles that stand in for
    //  dynamic register indexing cost more than the L1 hits they replace.)
    const int lane = lane_id();
    double prod = 1.0; int esum = 0;
    sg = 1.0;
#pragma unroll 1
    for (int c0 = 0; c0 < k; c0 += 4) {
        const int pw = k - c0 < 4 ? k - c0 : 4;
        const int rlo = want_inverse ? 0 : c0;
        // ---- panel
#pragma unroll 1
        for (int q = 0; q < pw; ++q) {
            const int c = c0 + q;
            double best = 0.0; int bi = c;
#pragma unroll 1
            for (int r = c + lane; r < k; r += W2B_WARP) { const double v = fabs(a[(size_t)r * ld + c]); if (v > best) { best = v; bi = r; } }
            double bv; const int p = wargmax_abs(best, bi, bv);
            if (!(bv > 0.0)) { sg = 0.0; lg = -1.7976931348623157e308; return; }
            if (lane == 0) wI[c] = p;
            if (p != c) {
#pragma unroll 1
                for (int j = lane; j < k; j += W2B_WARP) { const double t = a[(size_t)c * ld + j]; a[(size_t)c * ld + j] = a[(size_t)p * ld + j]; a[(size_t)p * ld + j] = t; }
                sg = -sg;
            }
            wsync();
            double pr[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) pr[t] = t < pw ? a[(size_t)c * ld + c0 + t] : 0.0;
            const double piv = q == 0 ? pr[0] : (q == 1 ? pr[1] : (q == 2 ? pr[2] : pr[3]));
            const double ipiv = 1.0 / piv;
            prod *= piv;
            {   // split the exponent off the running product (exact scaling)
#if W2B_DEVICE
                const int hi = __double2hiint(prod), e = ((hi >> 20) & 0x7ff) - 1023;
                prod = __hiloint2double((hi & 0x800fffff) | (1023 << 20), __double2loint(prod));
#else
                int e; prod = std::frexp(prod, &e) * 2.0; e -= 1;
#endif
                esum += e;
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) pr[t] = (t == q) ? ipiv : pr[t] * ipiv;      // scaled pivot row; its own column becomes the inverse column
            wsync();
#pragma unroll 1
            for (int r = rlo + lane; r < k; r += W2B_WARP) {
                double* row = a + (size_t)r * ld + c0;
                if (r == c) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) if (t < pw) row[t] = pr[t];
                } else {
                    const double f = row[q];
#pragma unroll
                    for (int t = 0; t < 4; ++t) if (t < pw) row[t] = (t == q) ? -f * ipiv : row[t] - f * pr[t];
                }
            }
            wsync();
        }
        // ---- every other column: one rank-4 update
        const int jlo = want_inverse ? 0 : (c0 + pw) & ~7;
        const int ilo = want_inverse ? 0 : (c0 + pw) & ~7;
#if W2B_DEVICE
        const int g = lane >> 2, t4 = lane & 3;
#pragma unroll 1
        for (int j0 = jlo; j0 < k; j0 += 8) {
            const int jb = j0 + g;                                                                  // B fragment: R[t4][j0 + g]
            const bool jb_ok = jb < k && t4 < pw && !(jb >= c0 && jb < c0 + pw);
            const double b = jb_ok ? a[(size_t)(c0 + t4) * ld + jb] : 0.0;
            const int jc = j0 + 2 * t4;                                                             // C fragment: rows i0 + g, columns jc, jc + 1
            const bool c0_ok = jc < k && !(jc >= c0 && jc < c0 + pw), c1_ok = jc + 1 < k && !(jc + 1 >= c0 && jc + 1 < c0 + pw);
            wsync();                                                                                // R is read before any row of this tile column is rewritten
#pragma unroll 2
            for (int i0 = ilo; i0 < k; i0 += 8) {
                const int ia = i0 + g;
                const bool in_b = ia >= c0 && ia < c0 + pw;
                const bool row_ok = ia < k && (want_inverse || ia >= c0 + pw);
                const double af = (ia < k && t4 < pw) ? a[(size_t)ia * ld + c0 + t4] : 0.0;        // A fragment: P[i0 + g][t4]
                double* cp = a + (size_t)ia * ld + jc;
                double d0 = (row_ok && c0_ok && !in_b) ? cp[0] : 0.0, d1 = (row_ok && c1_ok && !in_b) ? cp[1] : 0.0;
                dmma884(d0, d1, af, b);
                if (row_ok && c0_ok) cp[0] = d0;
                if (row_ok && c1_ok) cp[1] = d1;
            }
        }
        wsync();
#else
        for (int j = jlo; j < k; ++j) {
            if (j >= c0 && j < c0 + pw) continue;
            double rr[4];
            for (int t = 0; t < pw; ++t) rr[t] = a[(size_t)(c0 + t) * ld + j];
            for (int i = ilo; i < k; ++i) {
                if (!want_inverse && i < c0 + pw) continue;
                double v = (i >= c0 && i < c0 + pw) ? 0.0 : a[(size_t)i * ld + j];
                for (int t = 0; t < pw; ++t) v += a[(size_t)i * ld + c0 + t] * rr[t];
                a[(size_t)i * ld + j] = v;
            }
        }
#endif
    }
    lg = w_log(fabs(prod)) + (double)esum * 0.6931471805599453;
    if (prod < 0.0) sg = -sg;
    if (want_inverse) {
        // undo the row swaps as column swaps in reverse order: inv(P A) = inv(A) P^T
#pragma unroll 1
        for (int c = k - 1; c >= 0; --c) {
            const int p = wI[c];
            if (p != c) for (int r = lane; r < k; r += W2B_WARP) { const double t = a[(size_t)r * ld + c]; a[(size_t)r * ld + c] = a[(size_t)r * ld + p]; a[(size_t)r * ld + p] = t; }
            wsync();
        }
    }
}

// Rebuild every inverse hybridisation matrix from the flavour tau arrays (get_InvFullHybr Trace.F90:2279-2333):
// Delta(ia, ic) = F(tau_a[ia] - tau_c[ic]) ; M = Delta^{-1} is indexed (creator, annihilator).
// Builds directly the transposed system so that the in-place inverse lands in M's [creator][annihilator] layout.
W2B_HDN void rebuild_inverses(const Tables& T, const Chains& C, Ctx& x, const double* tcs, const double* tas, const int* kfs,
                              bool want_inverse, double& lg, double& sg) {
    lg = 0.0; sg = 1.0;
#pragma unroll 1
    for (int f = 0; f < T.nf; ++f) {
        int k = kfs[f];
        if (k == 0) continue;
        double* a = Mbuf(C, x, want_inverse ? f : T.nf);
        const double* tc = tcs + (size_t)f * C.kcap; const double* ta = tas + (size_t)f * C.kcap;
        // a[ia][ic] = Delta(ia, ic); its inverse inv[ic][ia] is what M stores at [creator ic][annihilator ia]
#pragma unroll 1
        for (int ia = 0; ia < k; ++ia)
#pragma unroll 1
            for (int ic = lane_id(); ic < k; ic += W2B_WARP) a[(size_t)ia * C.ld + ic] = hybrF(T, ta[ia] - tc[ic], f);
        wsync();
        double l, s;
        if (k >= LU_BLOCKED_MIN) invert_blocked(a, k, C.ld, x.wI, want_inverse, l, s);
        else invert_inplace(a, k, C.ld, x.wI, x.wQ, x.wR, want_inverse, l, s);
        lg += l; sg *= s;
    }
}

// ------------------------------------------------------------------ StepGlob (CTQMC.F90:4433-4554)
// Split in two halves around the local trace of the proposed configuration so that blocked traces can run in the trace
// kernel: prepare (update choice, relabelling, quantum numbers, determinant from scratch) returns true when a trace is
// needed; finish takes (log, sign) of that trace, draws the Metropolis number and applies the update.
W2B_HDN bool step_glob_prepare(const Tables& T, const Run& R, const Chains& C, Ctx& x, RareJob& J, bool with_lu = true) {
    log_begin(x, 3);
    bump(x, C_TRYGLOB);
    if (x.n == 0) { log_end(x, 2); return false; }
    const int nf = T.nf, nb = T.nb;
    int* gu = x.wI + C.kcap;               // gu[0] = c<->c^dag flip, gu[1..nf] = target flavour (bit index + 1)
    const double r = x.rng.next();
    int type;
    if (r < 0.3) {                         // gen_SpinFlipUpdate Trace.F90:1150-1166
        type = 3;
        if (lane_id() == 0) { gu[0] = 0; for (int i = 1; i <= nb; ++i) gu[i] = i + nb; for (int i = nb + 1; i <= nf; ++i) gu[i] = i - nb; }
    } else if (r < 0.6) {                  // gen_CAFlip 1186-1198
        type = 4;
        if (lane_id() == 0) { gu[0] = 1; for (int i = 1; i <= nf; ++i) gu[i] = i; }
    } else if (r < 0.8 && R.nsym > 0) {    // gen_SymUpdate 1170-1182: one of the user's symmetry moves or its inverse (CTQMC.F90:4463-4465)
        type = 2;
        const int isym = x.rng.randint(1, 2 * R.nsym) - 1;
        if (lane_id() == 0) { gu[0] = 0; for (int i = 1; i <= nf; ++i) gu[i] = ldg(&T.symmoves[isym * nf + i - 1]); }
    } else {                               // gen_GlobalUpdate_new 1103-1123
        type = 1;
        if (lane_id() == 0) for (int i = 1; i <= nf; ++i) gu[i] = i;
        wsync();
#pragma unroll 1
        for (int j = 1; j <= nf; ++j) { int i = x.rng.randint(j, nf); if (lane_id() == 0) { int t = gu[j]; gu[j] = gu[i]; gu[i] = t; } wsync(); }
        int flip = x.rng.randint(0, 1);
        if (lane_id() == 0) gu[0] = flip;
    }
    wsync();
    log_set(x, 4, type);
    const int flip = gu[0];
    // relabelled operator codes (globalUpdate Trace.F90:2126-2188); flavour f=2b+s <-> gu index b+nb*s+1
#pragma unroll 1
    for (int i = lane_id(); i < x.n; i += W2B_WARP) {
        int cd = x.code[i], f = cd >> 1, ty = cd & 1, b = f >> 1, s = f & 1;
        int g = gu[b + nb * s + 1] - 1;
        int f2 = (g % nb) * 2 + g / nb;
        x.sc_code[i] = (unsigned char)(f2 * 2 + (flip ? 1 - ty : ty));
    }
    wsync();
    double factor = 0.0; int new_outer = x.outer, new_state = R.statesampling ? x.ostate : -1;
    if (type == 3 && R.has_szt) {          // perform_spinflip_check_qn Trace.F90:1306-1339
        int st = ldg(&T.first_state[x.outer]);
        int dn = st & ((1 << nb) - 1), up = (st >> nb) & ((1 << nb) - 1);
        new_outer = ldg(&T.state2sst[up | (dn << nb)]);
        int fin = walk_store(T, x.sc_code, 0, x.n, new_outer, x.sc_sst, 0);
        factor = (fin == new_outer) ? 1.0 : 0.0;
        if (factor != 0.0 && R.statesampling) {                        // Trace.F90:1325-1332
            const int cand = new_outer; const double ot = outer_tau(T, x.tau, x.n);
            factor = wosc(T, x, ot, ot, x.outer, cand, nullptr, new_outer, new_state);
        }
    } else {                               // perform_global_check_qn Trace.F90:1345-1394
        int pold = 0, pnew = 0;
        // each lane walks one candidate outer superstate through the old and the new sequence
#pragma unroll 1
        for (int base = 0; base < T.nsst; base += W2B_WARP) {
            int k0 = base + lane_id(); bool act = k0 < T.nsst;
            int co = act ? k0 : -1, cn = act ? k0 : -1;
#pragma unroll 1
            for (int i = 0; i < x.n; ++i) {
                int a = x.code[i], b2 = x.sc_code[i];
                co = co >= 0 ? conn(T, co, a) : -1;
                cn = cn >= 0 ? conn(T, cn, b2) : -1;
            }
            bool oko = act && co == k0, okn = act && cn == k0;
            if (act) { x.wI[C.kcap + 32 + k0] = okn ? 1 : 0; x.wI[C.kcap + 32 + T.nsst + k0] = oko ? 1 : 0; }
            pold += wsum(oko ? 1 : 0); pnew += wsum(okn ? 1 : 0);
        }
        wsync();
        if (pold <= 0) { x.err |= ERR_GLOBAL_QN; log_end(x, 2); return false; }
        if (pnew > 0 && R.statesampling) {                             // Trace.F90:1372-1379: weighted choice among all states of the compatible superstates
            const double ot = outer_tau(T, x.tau, x.n);
            factor = wosc(T, x, ot, ot, -1, -1, x.wI + C.kcap + 32, new_outer, new_state);
            if (factor != 0.0) walk_store(T, x.sc_code, 0, x.n, new_outer, x.sc_sst, 0);
        } else if (pnew > 0) {
            factor = (double)pnew / (double)pold;
            int kk = x.rng.randint(1, pnew);
#pragma unroll 1
            for (int i = 0; i < T.nsst; ++i) if (x.wI[C.kcap + 32 + i]) { if (--kk <= 0) { new_outer = i; break; } }
            walk_store(T, x.sc_code, 0, x.n, new_outer, x.sc_sst, 0);
        }
    }
    if (factor == 0.0) { bump(x, C_QNGLOB); log_end(x, 2); return false; }
    long long sumk2 = 0; for (int f = 0; f < nf; ++f) sumk2 += (long long)x.ih[IH_KF + f] * x.ih[IH_KF + f];
    bump_bytes(x, 8ll * sumk2);                   // build Delta_new (write)
    // relabelled flavour tau arrays and orders
#pragma unroll 1
    for (int f = 0; f < nf; ++f) {
        int b = f >> 1, s = f & 1, g = gu[b + nb * s + 1] - 1, f2 = (g % nb) * 2 + g / nb;
        int k = x.ih[IH_KF + f];
        if (lane_id() == 0) J.knew[f2] = k;
        const double* oc = x.tc + (size_t)f * C.kcap; const double* oa = x.ta + (size_t)f * C.kcap;
        double* nc = x.sc_tc + (size_t)f2 * C.kcap; double* na = x.sc_ta + (size_t)f2 * C.kcap;
#pragma unroll 1
        for (int i = lane_id(); i < k; i += W2B_WARP) { nc[i] = flip ? oa[i] : oc[i]; na[i] = flip ? oc[i] : oa[i]; }
    }
    wsync();
    if (lane_id() == 0) { J.kind = 3; J.new_outer = new_outer; J.new_state = new_state; J.type = type; J.factor = factor; J.stored = 0; }
    wsync();
    if (!with_lu) return true;               // (callers that obtain the determinants elsewhere)
    // determinant of the new configuration from scratch (LU in the spare buffer).  The reference obtains the old determinant
    // from an LU of the stored inverses (get_FullLogDet, Trace.F90:2665-2685); the tracked Det is the same number up to the
    // rounding of the fast updates (pinned by the from-scratch tests), so the second LU is not repeated here.
    double dn_log, dn_sign;
    rebuild_inverses(T, C, x, x.sc_tc, x.sc_ta, J.knew, false, dn_log, dn_sign);
    if (lane_id() == 0) { J.dn_log = dn_log; J.dn_sign = dn_sign; }
    wsync();
    return true;
}
// finish, part a: Metropolis decision; on acceptance the relabelled configuration replaces the old one (the inverse matrices still
// have to be rebuilt: part b).  Returns true when accepted.
W2B_HDN bool step_glob_finish_a(const Tables& T, const Run& R, const Chains& C, Ctx& x, RareJob& J, double tn_log, double tn_sign) {
    const int nf = T.nf;
    const double dn_log = J.dn_log, dn_sign = J.dn_sign;
    const double dr_log = dn_log - x.det_log, dr_sign = dn_sign * x.det_sign;
    const double trr_log = tn_log - x.tr_log, trr_sign = tn_sign * x.tr_sign;
    const double prob = fabs(w_exp(trr_log + dr_log) * J.factor);
    const double rand = x.rng.next();
    log_set(x, 7, dr_log); log_set(x, 8, dr_sign); log_set(x, 9, trr_log); log_set(x, 10, trr_sign); log_set(x, 11, prob); log_set(x, 17, rand);
    if (!(rand < prob)) { log_end(x, 0); return false; }
    // ---- accept
    long long sumk2 = 0;
#pragma unroll 1
    for (int f = 0; f < nf; ++f) sumk2 += (long long)x.ih[IH_KF + f] * x.ih[IH_KF + f];
    bump_bytes(x, 8ll * sumk2);
#pragma unroll 1
    for (int i = lane_id(); i < x.n; i += W2B_WARP) { x.code[i] = x.sc_code[i]; x.sst[i] = x.sc_sst[i]; }
#pragma unroll 1
    for (int f = 0; f < nf; ++f) {
        int k = J.knew[f];
#pragma unroll 1
        for (int i = lane_id(); i < k; i += W2B_WARP) { x.tc[(size_t)f * C.kcap + i] = x.sc_tc[(size_t)f * C.kcap + i]; x.ta[(size_t)f * C.kcap + i] = x.sc_ta[(size_t)f * C.kcap + i]; }
    }
    wsync();
    if (lane_id() == 0) for (int f = 0; f < nf; ++f) x.ih[IH_KF + f] = J.knew[f];
    wsync();
    return true;
}
// finish, part b (the inverse matrices of the new configuration are in place): outer superstate, determinant, trace, signs
W2B_HDN void step_glob_finish_b(const Tables& T, const Run& R, const Chains& C, Ctx& x, RareJob& J, double tn_log, double tn_sign) {
    const int nf = T.nf;
    x.outer = J.new_outer;
    if (R.statesampling) { x.ostate = J.new_state; x.only = x.ostate; }
    x.det_log = J.dn_log; x.det_sign = J.dn_sign; x.tr_log = tn_log; x.tr_sign = tn_sign;
    if (T.dmax > 1) { commit_parts(T, x); slots_reset(C, x); }
    x.perm = perm_full(x.code, x.n);
    x.palt = 0;
#pragma unroll 1
    for (int f = 0; f < nf; ++f) x.palt |= palt_flavour(C, x, f) << f;
    update_cfgsign(T, x);
    bump(x, C_ACCGLOB);
    log_end(x, 1);
}
W2B_HDN void step_glob_finish(const Tables& T, const Run& R, const Chains& C, Ctx& x, RareJob& J, double tn_log, double tn_sign) {
    if (!step_glob_finish_a(T, R, C, x, J, tn_log, tn_sign)) return;
    double l2, s2;
    rebuild_inverses(T, C, x, x.tc, x.ta, x.ih + IH_KF, true, l2, s2);
    step_glob_finish_b(T, R, C, x, J, tn_log, tn_sign);
}

// ------------------------------------------------------------------ StepShiftTau (CTQMC.F90:4557-4835), same two halves
W2B_HDN bool step_shift_prepare(const Tables& T, const Run& R, const Chains& C, Ctx& x, RareJob& J) {
    log_begin(x, 4);
    if (x.n == 0) {                        // StepChangeOuter CTQMC.F90:4570-4638
        const double rand = x.rng.next();
        double acc = 0.0; int pick = T.nsst - 1;
        if (R.statesampling) {             // outer (superstate, eigenstate) ~ empty-trace value of the state (CTQMC.F90:4580-4608)
            int pstate = 0; bool done = false;
#pragma unroll 1
            for (int i = 0; i < T.nsst && !done; ++i) {
                const double* ev = T.evals + ldg(&T.offset[i]); const int d = ldg(&T.dim[i]);
#pragma unroll 1
                for (int j = 0; j < d; ++j) {
                    const double w = w_exp(-(ldg(&ev[j]) - T.egs) * T.beta) / T.empty_norm;
                    acc += w; pick = i; pstate = j;
                    if (acc >= rand && w != 0.0) { done = true; break; }
                }
            }
            x.outer = pick; x.ostate = pstate; x.only = pstate;
            x.tr_log = -(ldg(&T.evals[ldg(&T.offset[pick]) + pstate]) - T.egs) * T.beta; x.tr_sign = 1.0;
            if (T.dmax > 1) empty_parts(T, x, pick);
            update_cfgsign(T, x);
            log_set(x, 17, rand); log_end(x, 1);
            return false;
        }
#pragma unroll 1
        for (int i = 0; i < T.nsst; ++i) { double w = ldg(&T.empty_w[i]); acc += w; pick = i; if (acc >= rand && w != 0.0) break; }
        x.outer = pick; x.tr_log = ldg(&T.empty_log[pick]); x.tr_sign = 1.0;
        if (T.dmax > 1) empty_parts(T, x, pick);
        update_cfgsign(T, x);
        log_set(x, 17, rand); log_end(x, 1);
        return false;
    }
    const double dt = x.rng.next() * T.beta;
    log_set(x, 2, dt);
    if (dt == 0.0 || dt == T.beta) { log_end(x, 2); return false; }
    bump(x, C_TRYSHIFT);
    const int n = x.n;
    const int nwrap = count_lt(x.tau, n, dt);                     // operators passing through tau = 0
    const int first = nwrap == n ? 0 : nwrap;
    const int new_outer = first == 0 ? x.outer : (int)x.sst[first - 1];
#pragma unroll 1
    for (int i = lane_id(); i < n; i += W2B_WARP) {
        int src = first + i; if (src >= n) src -= n;
        double t = x.tau[src] - dt; if (t < 0.0) t += T.beta;     // modulo(tau - dt, beta)
        x.sc_tau[i] = t; x.sc_code[i] = x.code[src]; x.sc_sst[i] = x.sst[src];
    }
    wsync();
    int sst_new = new_outer, new_state = -1; double osf = 1.0;
    if (R.statesampling) osf = wosc(T, x, outer_tau(T, x.tau, n), outer_tau(T, x.sc_tau, n), x.outer, new_outer, nullptr, sst_new, new_state);   // CTQMC.F90:4702-4710
    if (lane_id() == 0) { J.kind = 4; J.dt = dt; J.new_outer = new_outer; J.new_state = new_state; J.factor = osf; J.stored = 0; }
    wsync();
    return true;
}
W2B_HDN void step_shift_finish(const Tables& T, const Run& R, const Chains& C, Ctx& x, RareJob& J, double tn_log, double tn_sign) {
    const double dt = J.dt;
    const int n = x.n, new_outer = J.new_outer;
    const int wp_old = x.wpos; const double wmin_old = x.wmin, wmax_old = x.wmax;
    const int sh = ((int)floor((dt / T.beta) * (double)(R.num_windows + 1))) % R.num_windows;
    x.wpos = ((wp_old - sh) % R.num_windows + R.num_windows) % R.num_windows;
    set_window_edges(T, R, x);
    const double osf = R.statesampling ? J.factor : 1.0;
    const double trr_log = osf == 0.0 ? 0.0 : tn_log - x.tr_log, trr_sign = osf == 0.0 ? 0.0 : tn_sign * x.tr_sign;
    const double prob = osf == 0.0 ? 0.0 : fabs(osf * (trr_sign * w_exp(trr_log)));
    const double rand = x.rng.next();
    log_set(x, 9, trr_log); log_set(x, 10, trr_sign); log_set(x, 11, prob); log_set(x, 17, rand);
    if (!(rand < prob)) {
        x.wpos = wp_old; x.wmin = wmin_old; x.wmax = wmax_old;
        if (J.stored) slots_reset(C, x);                          // the trace kernel overwrote the caches with the rejected configuration
        log_end(x, 0); return;
    }
    // ---- accept: update_InvFullHybrShiftTau_diag Trace.F90:2355-2457 (cyclic roll + negation of wrapped rows/columns)
    double detrat = 1.0;
#pragma unroll 1
    for (int f = 0; f < T.nf; ++f) {
        const int k = x.ih[IH_KF + f];
        if (k == 0) continue;
        double* tc = x.tc + (size_t)f * C.kcap; double* ta = x.ta + (size_t)f * C.kcap;
        const int w1 = count_lt(tc, k, dt), w2 = count_lt(ta, k, dt);
        if (!((w1 == 0 && w2 == 0) || (w1 == k && w2 == k))) {
            const double* src = Mbuf(C, x, f); double* dst = Mbuf(C, x, T.nf);
            bump_bytes(x, 16ll * k * k);
#pragma unroll 1
            for (int i = 0; i < k; ++i) {
                int in = i - w1; double si = 1.0; if (in < 0) { in += k; si = -1.0; }
#pragma unroll 1
                for (int j = lane_id(); j < k; j += W2B_WARP) {
                    int jn = j - w2; double sj = si; if (jn < 0) { jn += k; sj = -si; }
                    dst[(size_t)in * C.ld + jn] = sj * src[(size_t)i * C.ld + j];
                }
            }
            wsync();
            if (lane_id() == 0) { int t = x.ih[IH_MSLOT + f]; x.ih[IH_MSLOT + f] = x.ih[IH_MSLOT + T.nf]; x.ih[IH_MSLOT + T.nf] = t; }
            wsync();
        }
        if ((k & 1) && ((w1 + w2) & 1)) detrat = -detrat;
        // rotate the flavour tau arrays
        double* scc = x.sc_tc + (size_t)f * C.kcap; double* sca = x.sc_ta + (size_t)f * C.kcap;
#pragma unroll 1
        for (int i = lane_id(); i < k; i += W2B_WARP) {
            int s1 = i + w1; if (s1 >= k) s1 -= k;
            int s2 = i + w2; if (s2 >= k) s2 -= k;
            double v1 = tc[s1] - dt; if (v1 < 0.0) v1 += T.beta;
            double v2 = ta[s2] - dt; if (v2 < 0.0) v2 += T.beta;
            scc[i] = v1; sca[i] = v2;
        }
        wsync();
#pragma unroll 1
        for (int i = lane_id(); i < k; i += W2B_WARP) { tc[i] = scc[i]; ta[i] = sca[i]; }
        wsync();
    }
#pragma unroll 1
    for (int i = lane_id(); i < n; i += W2B_WARP) { x.tau[i] = x.sc_tau[i]; x.code[i] = x.sc_code[i]; x.sst[i] = x.sc_sst[i]; }
    wsync();
    x.outer = new_outer;
    if (R.statesampling) { x.ostate = J.new_state; x.only = x.ostate; }
    x.tr_log = tn_log; x.tr_sign = tn_sign; x.det_sign *= detrat;
    if (T.dmax > 1) { commit_parts(T, x); slots_reset(C, x); if (J.stored) x.kv = n; }   // kets of the shifted sequence are already cached
    x.npairs = -1;
    x.perm = perm_full(x.code, x.n);
    x.palt = 0; for (int f = 0; f < T.nf; ++f) x.palt |= palt_flavour(C, x, f) << f;
    update_cfgsign(T, x);
    bump(x, C_ACCSHIFT);
    log_end(x, 1);
}
// both halves back to back with the trace in between (1x1 blocks; host paths that do not use the trace kernel)
W2B_HD void step_rare(const Tables& T, const Run& R, const Chains& C, Ctx& x, RareJob& J, int kind) {
    const bool go = kind == 3 ? step_glob_prepare(T, R, C, x, J) : step_shift_prepare(T, R, C, x, J);
    if (!go) return;
    const double* tau = kind == 3 ? x.tau : x.sc_tau;
    double tn_log, tn_sign;
    if (T.dmax > 1) full_trace(T, x, tau, x.sc_code, x.n, J.new_outer, tn_log, tn_sign, J.new_state);
    else trace_full_scalar(T, tau, x.sc_code, x.sc_sst, x.n, J.new_outer, tn_log, tn_sign);
    if (kind == 3) step_glob_finish(T, R, C, x, J, tn_log, tn_sign); else step_shift_finish(T, R, C, x, J, tn_log, tn_sign);
}

// ------------------------------------------------------------------ StepAdd / StepRem with change_outer (PercentageOuterMove)
// Served like global moves: the proposed sequence is staged in the scratch arrays and its local trace evaluated from scratch.
// Stages the proposed sequence of an outer pair move: times and codes with the pair inserted / taken out, superstates walked from
// the new outer superstate P.begin_sst; 1x1 blocks: also the trace ratio (P.dlog, P.dsign).  Returns false when it does not fit.
W2B_HDN bool outer_stage(const Tables& T, const Run& R, const Chains& C, Ctx& x, Prop& P, RareJob& J, bool add) {
    const int n = x.n, i1 = P.i1, i2 = P.i2, nn = add ? n + 2 : n - 2, new_outer = P.begin_sst;
    if (nn > C.ops_cap) { x.err |= ERR_CAPACITY; log_end(x, 0); return false; }
    const int codeA = P.f * 2 + P.type1, codeB = P.f * 2 + (1 - P.type1);
#pragma unroll 1
    for (int i = lane_id(); i < nn; i += W2B_WARP) {
        int src;                                                     // >= 0: old operator; -1 / -2: the new pair
        if (add) src = i < i1 ? i : (i == i1 ? -1 : (i <= i2 ? i - 1 : (i == i2 + 1 ? -2 : i - 2)));
        else src = i < i1 ? i : (i < i2 - 1 ? i + 1 : i + 2);
        x.sc_tau[i] = src >= 0 ? x.tau[src] : (src == -1 ? P.t1 : P.t2);
        x.sc_code[i] = src >= 0 ? x.code[src] : (unsigned char)(src == -1 ? codeA : codeB);
    }
    wsync();
    walk_store(T, x.sc_code, 0, nn, new_outer, x.sc_sst, 0);
    int new_state = -1;
    if (R.statesampling) {                 // the new outer eigenstate is drawn inside the new outer superstate (CTQMC.F90:3471-3480, 3979-3996)
        int sn; double osf = wosc(T, x, outer_tau(T, x.tau, n), outer_tau(T, x.sc_tau, nn), x.outer, new_outer, nullptr, sn, new_state);
        if (lane_id() == 0) { P.factor = mul_rn(osf, P.factor); P.new_state = new_state; }
    }
    if (T.dmax == 1) {
        double tl, ts;
        trace_full_scalar(T, x.sc_tau, x.sc_code, x.sc_sst, nn, new_outer, tl, ts);
        if (lane_id() == 0) { P.dlog = tl - x.tr_log; P.dsign = ts * x.tr_sign; }
    }
    if (lane_id() == 0) { J.kind = add ? 5 : 6; J.new_outer = new_outer; J.new_state = new_state; J.nnew = nn; J.stored = 0; }
    wsync();
    return true;
}
// first half (kinds 5 / 6); returns true when the proposal survived and is staged
W2B_HDN bool outer_prepare(const Tables& T, const Run& R, const Chains& C, Ctx& x, Prop& P, RareJob& J, int kind) {
    const bool go = kind == 5 ? add_propose<true>(T, R, C, x, P) : rem_propose<true>(T, R, C, x, P);
    return go && outer_stage(T, R, C, x, P, J, kind == 5);
}
// the whole move on one warp (1x1 blocks; host paths that do not use the trace kernel)
W2B_HDN void step_outer(const Tables& T, const Run& R, const Chains& C, const Accum& A, Ctx& x, Prop& P, RareJob& J, int kind) {
    if (!outer_prepare(T, R, C, x, P, J, kind)) return;
    if (T.dmax > 1) { double tl, ts; full_trace(T, x, x.sc_tau, x.sc_code, J.nnew, J.new_outer, tl, ts, J.new_state); }     // parts in x.mpt
    if (kind == 5) add_decide<true>(T, R, C, A, x, P); else rem_decide<true>(T, R, C, A, x, P);
}

// ------------------------------------------------------------------ 4-operator moves (Percentage4OperatorMove, diagonal hybridisation)
// StepAdd4 / StepRem4 (CTQMC.F90:3615-3889, 4153-4430): two pairs are inserted / removed at once.  The reference takes the determinant
// of every proposed configuration from scratch (get_InvFullHybr, CTQMC.F90:3816); here the RATIO is a rank-2 form on the stored
// inverses (four_det_ratio) and only accepted moves rebuild the inverse matrices (blocked tensor-pipe elimination); the proposal
// ratio sums over the orders in which the inverse move could take the four operators out (rempair_factor).  Served like global moves (kinds 7 / 8 of the rare machinery): proposed sequence staged
// in the scratch arrays, whole-sequence local trace.  All pair counts are taken over ONE time-sorted list U (the stored
// configuration, for insertions with the four new operators merged in) with some of the four "special" operators masked out.
W2B_HD bool four_masked(int i, const int* sp, int mask) {
    return ((mask & 1) && i == sp[0]) || ((mask & 2) && i == sp[1]) || ((mask & 4) && i == sp[2]) || ((mask & 8) && i == sp[3]);
}
// partners of operator i in gen_OperRemove's enumeration (Trace.F90:4886-4914): later operators of the same flavour and opposite type
// with tau <= wmax and tau <= tau_i + taudiff_max; `want` > 0: returns the index of the want-th partner instead of the count
W2B_HDN int four_partners(const Run& R, const Ctx& x, const double* tau, const unsigned char* code, int n, const int* sp, int mask, int i, int want) {
    const int cd = code[i]; const double ti = tau[i];
    int c = 0;
#pragma unroll 1
    for (int m = i + 1; m < n; ++m) {
        const int cm = code[m];
        if ((cm >> 1) != (cd >> 1) || four_masked(m, sp, mask)) continue;
        const double v = tau[m];
        if (v > x.wmax || v > ti + R.taudiff_max) break;
        if ((cm ^ cd) & 1) { ++c; if (c == want) return m; }
    }
    return want > 0 ? -1 : c;
}
// NPairs of U minus the masked specials (ensure_valid_NPairs_diag Trace.F90:1770-1815); r > 0: the r-th pair in enumeration order -> (i1, i2)
W2B_HDN int four_pairs(const Run& R, const Ctx& x, const double* tau, const unsigned char* code, int n, const int* sp, int mask, int r, int& i1, int& i2) {
    const int ia = count_lt(tau, n, x.wmin), ib = count_le(tau, n, x.wmax);
    int cum = 0;
#pragma unroll 1
    for (int base = ia; base < ib; base += W2B_WARP) {
        const int i = base + lane_id();
        const int c = (i < ib && !four_masked(i, sp, mask)) ? four_partners(R, x, tau, code, n, sp, mask, i, 0) : 0;
        const int incl = wscan_incl(c), total = wbcast(incl, W2B_WARP - 1);
        if (r > 0 && cum + total >= r) {
            const unsigned hit = wballot(cum + incl >= r);
            int src = 0;
#if W2B_DEVICE
            src = __ffs(hit) - 1;
#endif
            i1 = base + src;
            i2 = four_partners(R, x, tau, code, n, sp, mask, i1, r - (cum + wbcast(incl - c, src)));
            return cum + total;
        }
        cum += total;
    }
    return cum;
}
W2B_HD bool is_rempair(const Run& R, const Ctx& x, double t1, int c1, double t2, int c2) {      // is_rempair_diag Trace.F90:2005-2030
    return (c1 >> 1) == (c2 >> 1) && ((c1 ^ c2) & 1) && t1 >= x.wmin && t1 <= x.wmax && t2 >= x.wmin && t2 <= x.wmax && fabs(t1 - t2) <= R.taudiff_max;
}
// Determinant ratio of a 4-operator move WITHOUT an elimination: Woodbury / Jacobi rank-2 forms on the stored inverse matrices.
//   two pairs of different flavours: the product of the two rank-1 ratios (S = F - R M Q, resp. the element M(c, a));
//   insertion into ONE flavour:  det Delta' / det Delta = (-1)^(i1+i2+j1+j2) det S,  S_uv = F(a_u - c_v) - R_u M Q_v  (2 x 2 Schur complement;
//     rows u = the new annihilators, columns v = the new creators, both in time order; i, j their positions in the new matrix);
//   removal from ONE flavour:    det Delta' / det Delta = (-1)^(i1+i2+j1+j2) det [ M(c_v, a_u) ]   (Jacobi's complementary-minor identity).
// O(k^2) per proposal instead of the reference's from-scratch determinant of every flavour (CTQMC.F90:3816-3817); accepted moves still
// rebuild the inverse matrices from scratch as the reference does, which also resets the rounding of the fast updates.
// t / code: the four operators; returns (log|ratio|, sign); sign 0 = singular.
W2B_HDN void four_det_ratio(const Tables& T, const Chains& C, Ctx& x, bool add, const double* t, const int* code, double& rlog, double& rsign) {
    double ratio = 1.0; int par = 0;
    const int fA = code[0] >> 1, fB = code[2] >> 1;
    // creator / annihilator time of each pair (each pair holds one of either)
    double tc2[2], ta2[2];
#pragma unroll 1
    for (int pr = 0; pr < 2; ++pr) { const bool first_is_a = code[2 * pr] & 1; ta2[pr] = first_is_a ? t[2 * pr] : t[2 * pr + 1]; tc2[pr] = first_is_a ? t[2 * pr + 1] : t[2 * pr]; }
    if (fA != fB) {
#pragma unroll 1
        for (int pr = 0; pr < 2; ++pr) {
            const int f = pr == 0 ? fA : fB, k = x.ih[IH_KF + f];
            const double* tc = x.tc + (size_t)f * C.kcap; const double* ta = x.ta + (size_t)f * C.kcap;
            const int p1 = count_lt(tc, k, tc2[pr]), p2 = count_lt(ta, k, ta2[pr]);
            ratio *= add ? det_ratio_add(T, C, x, f, tc2[pr], ta2[pr]) : Mbuf(C, x, f)[(size_t)p1 * C.ld + p2];
            par ^= (p1 - p2) & 1;
        }
    } else {
        const int f = fA, k = x.ih[IH_KF + f];
        const double* tc = x.tc + (size_t)f * C.kcap; const double* ta = x.ta + (size_t)f * C.kcap;
        const double c0 = fmin(tc2[0], tc2[1]), c1 = fmax(tc2[0], tc2[1]), a0 = fmin(ta2[0], ta2[1]), a1 = fmax(ta2[0], ta2[1]);
        const int jc0 = count_lt(tc, k, c0), jc1 = count_lt(tc, k, c1), ia0 = count_lt(ta, k, a0), ia1 = count_lt(ta, k, a1);
        if (add) {
            // positions in the NEW arrays: jc0, jc1 + 1, ia0, ia1 + 1 -> the two +1 cancel in the parity
            par = (jc0 + jc1 + ia0 + ia1) & 1;
            const double S00 = det_ratio_add(T, C, x, f, c0, a0);               // leaves V0 = R(a0)^T M in wV
            double part = 0.0;
#pragma unroll 1
            for (int j = lane_id(); j < k; j += W2B_WARP) part += x.wV[j] * hybrF(T, ta[j] - c1, f);
            const double S01 = hybrF(T, a0 - c1, f) - wsum(part);
            wsync();
            const double S11 = det_ratio_add(T, C, x, f, c1, a1);               // V1 = R(a1)^T M
            part = 0.0;
#pragma unroll 1
            for (int j = lane_id(); j < k; j += W2B_WARP) part += x.wV[j] * hybrF(T, ta[j] - c0, f);
            const double S10 = hybrF(T, a1 - c0, f) - wsum(part);
            wsync();
            ratio = S00 * S11 - S01 * S10;
        } else {
            par = (jc0 + jc1 + ia0 + ia1) & 1;
            const double* M = Mbuf(C, x, f); const int ld = C.ld;
            ratio = M[(size_t)jc0 * ld + ia0] * M[(size_t)jc1 * ld + ia1] - M[(size_t)jc0 * ld + ia1] * M[(size_t)jc1 * ld + ia0];
        }
    }
    if (ratio == 0.0) { rlog = -1.7976931348623157e308; rsign = 0.0; return; }
    rlog = w_log(fabs(ratio)); rsign = sgn1(ratio) * (par ? -1.0 : 1.0);
}

// first half; returns true when a proposal is staged: sequence, flavour time arrays and orders in the scratch arrays, determinant in
// J.dn_*, proposal factor in J.factor, new pair count in J.np_new
W2B_HDN bool four_prepare(const Tables& T, const Run& R, const Chains& C, Ctx& x, RareJob& J, int kind) {
    const bool add = kind == 7;
    const int n = x.n, nf = T.nf;
    log_begin(x, kind);
    bump(x, add ? C_TRYADD4 : C_TRYREM4);
    if (x.npairs < 0) x.npairs = count_pairs_full(T, R, C, x);        // ensure_valid_NPairs CTQMC.F90:3651 / gen_OperRemove
    const int N0 = x.npairs;
    const double W = x.wmax - x.wmin;
    double t[4]; int code[4]; int* sp = x.wI;                         // sp[0..3]: list indices of the four special operators in U
    const double* Ut; const unsigned char* Uc; int nU, nn;
    double fac, rf; int np_new, N12, N34, dummy1, dummy2;
    if (add) {
        nn = n + 4; nU = nn;
        if (nn > C.ops_cap) { x.err |= ERR_CAPACITY; log_end(x, 0); return false; }
        fac = 1.0;                                                    // pair_OperAdd with N = 4, force_diagonal (Trace.F90:3657-3698)
#pragma unroll 1
        for (int pr = 0; pr < 2; ++pr) {
            const double tau_a = add_rn(x.wmin, mul_rn(W, x.rng.next()));
            const int b = x.rng.randint(1, T.nb) - 1, sp_ = x.rng.randint(1, 2) - 1;
            const double tmin = fmax(tau_a - R.taudiff_max, x.wmin), w2 = fmin(tau_a + R.taudiff_max, x.wmax) - tmin;
            const double tau_c = add_rn(tmin, mul_rn(w2, x.rng.next()));
            fac = mul_rn(mul_rn(fac, W), w2);
            const int f = b * 2 + sp_, lo = 2 * pr, hi = 2 * pr + 1;    // each pair in time order (process_OperAdd)
            if (tau_a < tau_c) { t[lo] = tau_a; code[lo] = f * 2 + 1; t[hi] = tau_c; code[hi] = f * 2; }
            else { t[lo] = tau_c; code[lo] = f * 2; t[hi] = tau_a; code[hi] = f * 2 + 1; }
        }
        log_set(x, 2, fmin(t[0], t[2])); log_set(x, 3, fmax(t[1], t[3])); log_set(x, 4, (code[0] >> 1) * nf + (code[2] >> 1));
        bool equal = false;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) { bool e; count_lt(x.tau, n, t[j], e); equal |= e; for (int j2 = 0; j2 < j; ++j2) equal |= t[j2] == t[j]; }
        if (equal) { log_end(x, 2); return false; }
        // U = the proposed sequence: stored operators shifted by the number of new ones before them, new ones at their ranks
#pragma unroll 1
        for (int i = lane_id(); i < n; i += W2B_WARP) {
            const double v = x.tau[i]; const int sh = (t[0] < v) + (t[1] < v) + (t[2] < v) + (t[3] < v);
            x.sc_tau[i + sh] = v; x.sc_code[i + sh] = x.code[i];
        }
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            int pos = count_lt(x.tau, n, t[j]);
#pragma unroll 1
            for (int j2 = 0; j2 < 4; ++j2) pos += (j2 != j && t[j2] < t[j]) ? 1 : 0;
            if (lane_id() == 0) { x.sc_tau[pos] = t[j]; x.sc_code[pos] = (unsigned char)code[j]; sp[j] = pos; }
        }
        wsync();
        Ut = x.sc_tau; Uc = x.sc_code;
        N12 = four_pairs(R, x, Ut, Uc, nU, sp, 4 | 8, 0, dummy1, dummy2);
        N34 = four_pairs(R, x, Ut, Uc, nU, sp, 1 | 2, 0, dummy1, dummy2);
        np_new = four_pairs(R, x, Ut, Uc, nU, sp, 0, 0, dummy1, dummy2);
    } else {
        nn = n - 4; nU = n; Ut = x.tau; Uc = x.code;
        log_set(x, 14, N0);
        if (N0 <= 0) { log_end(x, 2); return false; }
        // gen_OperRemove twice: the second pair is drawn from the configuration without the first
        int i1 = -1, i2 = -1;
        if (lane_id() == 0) { sp[0] = sp[1] = sp[2] = sp[3] = -1; }
        wsync();
        four_pairs(R, x, Ut, Uc, nU, sp, 0, x.rng.randint(1, N0), i1, i2);
        if (i1 < 0 || i2 < 0) { x.err |= ERR_PAIR_LOGIC; log_end(x, 2); return false; }
        if (lane_id() == 0) { sp[0] = i1; sp[1] = i2; }
        wsync();
        N34 = four_pairs(R, x, Ut, Uc, nU, sp, 1 | 2, 0, dummy1, dummy2);
        if (N34 <= 0) { log_end(x, 2); return false; }
        int i3 = -1, i4 = -1;
        four_pairs(R, x, Ut, Uc, nU, sp, 1 | 2, x.rng.randint(1, N34), i3, i4);
        if (i3 < 0 || i4 < 0) { x.err |= ERR_PAIR_LOGIC; log_end(x, 2); return false; }
        if (lane_id() == 0) { sp[2] = i3; sp[3] = i4; }
        wsync();
#pragma unroll 1
        for (int j = 0; j < 4; ++j) { t[j] = Ut[sp[j]]; code[j] = Uc[sp[j]]; }
        // inverse tau-choice probability of the insertion: (W w2(annihilator 1)) (W w2(annihilator 2))   (CTQMC.F90:4194-4195)
        fac = 1.0;
#pragma unroll 1
        for (int pr = 0; pr < 2; ++pr) {
            const double tan = (code[2 * pr] & 1) ? t[2 * pr] : t[2 * pr + 1];
            fac *= W * (fmin(tan + R.taudiff_max, x.wmax) - fmax(tan - R.taudiff_max, x.wmin));
        }
        N12 = four_pairs(R, x, Ut, Uc, nU, sp, 4 | 8, 0, dummy1, dummy2);
        np_new = four_pairs(R, x, Ut, Uc, nU, sp, 15, 0, dummy1, dummy2);
        log_set(x, 2, fmin(t[0], t[2])); log_set(x, 3, fmax(t[1], t[3])); log_set(x, 4, (code[0] >> 1) * nf + (code[2] >> 1));
    }
    {   // rempair_factor: the four operators can leave in two or in four orders (CTQMC.F90:3677-3766, 4204-4290)
        const double Nref = add ? (double)np_new : (double)N0;
        int N14 = 0, N23 = 0; bool four_ways = false;
        if (is_rempair(R, x, t[0], code[0], t[3], code[3]) && is_rempair(R, x, t[1], code[1], t[2], code[2])) {
            four_ways = true;
            N23 = four_pairs(R, x, Ut, Uc, nU, sp, 1 | 8, 0, dummy1, dummy2); N14 = four_pairs(R, x, Ut, Uc, nU, sp, 2 | 4, 0, dummy1, dummy2);
        } else if (is_rempair(R, x, t[0], code[0], t[2], code[2]) && is_rempair(R, x, t[1], code[1], t[3], code[3])) {
            four_ways = true;
            N23 = four_pairs(R, x, Ut, Uc, nU, sp, 1 | 4, 0, dummy1, dummy2); N14 = four_pairs(R, x, Ut, Uc, nU, sp, 2 | 8, 0, dummy1, dummy2);
        }
        if (four_ways) { rf = (1.0 / Nref) * (1.0 / (double)N12 + 1.0 / (double)N34 + 1.0 / (double)N14 + 1.0 / (double)N23); fac /= 4.0; }
        else { rf = (1.0 / Nref) * (1.0 / (double)N12 + 1.0 / (double)N34); fac /= 2.0; }
    }
    const double nb2 = (double)((2 * T.nb) * (2 * T.nb));
    const double factor = add ? fac * nb2 * rf : 1.0 / nb2 / fac / rf;
    if (!add) {                                                       // proposed sequence = U without the four
#pragma unroll 1
        for (int i = lane_id(); i < n; i += W2B_WARP) {
            if (four_masked(i, sp, 15)) continue;
            const int sh = (sp[0] < i) + (sp[1] < i) + (sp[2] < i) + (sp[3] < i);
            x.sc_tau[i - sh] = x.tau[i]; x.sc_code[i - sh] = x.code[i];
        }
        wsync();
    }
    // quantum numbers: the proposed sequence has to close on the unchanged outer superstate (check_sst_sequence over the changed
    // sub-string, CTQMC.F90:3769-3796; outside it nothing changes and the superstate maps are injective)
    if (walk_store(T, x.sc_code, 0, nn, x.outer, x.sc_sst, 0) != x.outer) { log_end(x, 2); return false; }
    // flavour time arrays of the proposed sequence: rank among the operators of the same code
#pragma unroll 1
    for (int c = 0; c < 2 * nf; ++c) {
        double* dst = ((c & 1) ? x.sc_ta : x.sc_tc) + (size_t)(c >> 1) * C.kcap;
        int run = 0;
#pragma unroll 1
        for (int base = 0; base < nn; base += W2B_WARP) {
            const int i = base + lane_id(); const bool mine = i < nn && (int)x.sc_code[i] == c;
            const int incl = wscan_incl(mine ? 1 : 0);
            if (mine) { if (run + incl - 1 < C.kcap) dst[run + incl - 1] = x.sc_tau[i]; }
            run += wbcast(incl, W2B_WARP - 1);
        }
        if (run > C.kcap) { x.err |= ERR_CAPACITY; log_end(x, 0); return false; }
        if (!(c & 1) && lane_id() == 0) J.knew[c >> 1] = run;
    }
    wsync();
    double dn_log, dn_sign;
    {   // determinant of the proposed configuration = stored determinant x rank-2 ratio (no elimination per proposal)
        double rl, rs;
        four_det_ratio(T, C, x, add, t, code, rl, rs);
        dn_log = rs == 0.0 ? rl : x.det_log + rl; dn_sign = x.det_sign * rs;
    }
    if (lane_id() == 0) { J.kind = kind; J.new_outer = x.outer; J.new_state = x.only; J.nnew = nn; J.stored = 0; J.factor = factor; J.np_new = np_new; J.dn_log = dn_log; J.dn_sign = dn_sign; }
    wsync();
    return true;
}
// second half: Metropolis; on acceptance the staged configuration replaces the stored one and the inverse matrices are rebuilt
// Rank-2 UPDATE of the inverse matrices for an accepted 4-operator move (the reference re-inverts every flavour from scratch,
// CTQMC.F90:3856-3866): the block (Woodbury) update of a flavour that gains / loses two rows and columns is applied in factored form,
// as two successive bordering / Schur-complement steps on the stored inverse (det_ratio_add + mat_add, or mat_rem on the pivot), O(k^2)
// instead of O(nf k^3); a flavour that gains / loses one pair takes one step.  The pairs are taken in time order (earliest new creator
// with earliest new annihilator).  Returns false -- the caller then re-inverts from scratch -- when an intermediate pivot is too small
// for the factored form (|S| < 1e-6: the 2 x 2 Schur complement can be regular while its leading element vanishes).
// On success x.tc / x.ta hold the proposed flavour arrays (x.sc_tc / x.sc_ta) and IH_KF the proposed orders.
W2B_HDN bool four_update_rank2(const Tables& T, const Chains& C, Ctx& x, const RareJob& J) {
#pragma unroll 1
    for (int f = 0; f < T.nf; ++f) {
        const int knew = J.knew[f];
        double* tc = x.tc + (size_t)f * C.kcap; double* ta = x.ta + (size_t)f * C.kcap;
        const double* nc = x.sc_tc + (size_t)f * C.kcap; const double* na = x.sc_ta + (size_t)f * C.kcap;
#pragma unroll 1
        while (x.ih[IH_KF + f] < knew) {                          // one bordering step: the earliest missing creator / annihilator
            const int k = x.ih[IH_KF + f];
            int p1 = 0, p2 = 0;
            while (p1 < k && tc[p1] == nc[p1]) ++p1;
            while (p2 < k && ta[p2] == na[p2]) ++p2;
            const double tau_c = nc[p1], tau_a = na[p2];
            const double S = det_ratio_add(T, C, x, f, tau_c, tau_a);
            if (!(fabs(S) >= 1e-6)) return false;
            bump_bytes(x, 8ll * (2 * k * k + (k + 1) * (k + 1)));
            mat_add(C, x, f, T.nf, p1, p2, S);
            shift_up(tc, p1, k, 1); shift_up(ta, p2, k, 1);
            if (lane_id() == 0) { tc[p1] = tau_c; ta[p2] = tau_a; }
            wsync();
        }
#pragma unroll 1
        while (x.ih[IH_KF + f] > knew) {                          // one Schur-complement step: the earliest surplus creator / annihilator
            const int k = x.ih[IH_KF + f];
            int p1 = 0, p2 = 0;
            while (p1 < k - 1 && tc[p1] == nc[p1]) ++p1;
            while (p2 < k - 1 && ta[p2] == na[p2]) ++p2;
            const double piv = Mbuf(C, x, f)[(size_t)p1 * C.ld + p2];
            if (!(fabs(piv) >= 1e-6)) return false;
            bump_bytes(x, 8ll * (k * k + (k - 1) * (k - 1)));
            mat_rem(C, x, f, T.nf, p1, p2);
            shift_down(tc, p1, k, 1); shift_down(ta, p2, k, 1);
            wsync();
        }
    }
    return true;
}
W2B_HDN void four_finish(const Tables& T, const Run& R, const Chains& C, Ctx& x, RareJob& J, double tn_log, double tn_sign) {
    const int nf = T.nf, nn = J.nnew;
    const double dr_log = J.dn_log - x.det_log, dr_sign = J.dn_sign * x.det_sign;
    const double trr_log = tn_log - x.tr_log, trr_sign = tn_sign * x.tr_sign;
    const double rand = x.rng.next();
    const double prob = fabs(J.factor * w_exp(trr_log + dr_log));
    log_set(x, 7, dr_log); log_set(x, 8, dr_sign); log_set(x, 9, trr_log); log_set(x, 10, trr_sign); log_set(x, 11, prob); log_set(x, 14, J.np_new); log_set(x, 17, rand);
    if (!(rand < prob)) { log_end(x, 0); return; }
#pragma unroll 1
    for (int i = lane_id(); i < nn; i += W2B_WARP) { x.tau[i] = x.sc_tau[i]; x.code[i] = x.sc_code[i]; x.sst[i] = x.sc_sst[i]; }
    wsync();
    x.n = nn;
    if (R.four_rank2 && four_update_rank2(T, C, x, J)) {                     // rank-2 update of the stored inverses; the determinant follows the rank-2 ratio
        x.det_log = J.dn_log; x.det_sign = J.dn_sign;
        bump(x, C_FOUR_RANK2);
    } else {
#pragma unroll 1
        for (int f = 0; f < nf; ++f) {
            const int k = J.knew[f];
#pragma unroll 1
            for (int i = lane_id(); i < k; i += W2B_WARP) { x.tc[(size_t)f * C.kcap + i] = x.sc_tc[(size_t)f * C.kcap + i]; x.ta[(size_t)f * C.kcap + i] = x.sc_ta[(size_t)f * C.kcap + i]; }
        }
        wsync();
        if (lane_id() == 0) for (int f = 0; f < nf; ++f) x.ih[IH_KF + f] = J.knew[f];
        wsync();
        double l2, s2;
        rebuild_inverses(T, C, x, x.tc, x.ta, x.ih + IH_KF, true, l2, s2);   // get_InvFullHybr: inverse matrices AND determinant from scratch (CTQMC.F90:3856-3866)
        x.det_log = l2; x.det_sign = s2;
    }
    x.tr_log = tn_log; x.tr_sign = tn_sign;
    if (T.dmax > 1) { commit_parts(T, x); slots_reset(C, x); }
    x.npairs = J.np_new;
    x.perm = perm_full(x.code, x.n);
    x.palt = 0;
#pragma unroll 1
    for (int f = 0; f < nf; ++f) x.palt |= palt_flavour(C, x, f) << f;
    update_cfgsign(T, x);
    bump(x, J.kind == 7 ? C_ACCADD4 : C_ACCREM4);
    log_end(x, 1);
}
// the whole move on one warp (1x1 blocks; host paths that do not use the trace kernel)
W2B_HDN void step_four(const Tables& T, const Run& R, const Chains& C, Ctx& x, RareJob& J, int kind) {
    if (!four_prepare(T, R, C, x, J, kind)) return;
    double tn_log, tn_sign;
    if (T.dmax > 1) full_trace(T, x, x.sc_tau, x.sc_code, J.nnew, x.outer, tn_log, tn_sign, x.only);
    else trace_full_scalar(T, x.sc_tau, x.sc_code, x.sc_sst, J.nnew, x.outer, tn_log, tn_sign);
    four_finish(T, R, C, x, J, tn_log, tn_sign);
}

// ------------------------------------------------------------------ chain load / store
W2B_HDN void ctx_bind(const Tables& T, const Run& R, const Chains& C, int chain, Ctx& x, double* wscratch, int* iscratch, double* mscratch) {
    x.tau = C.op_tau + (size_t)chain * C.ops_cap; x.code = C.op_code + (size_t)chain * C.ops_cap; x.sst = C.op_sst + (size_t)chain * C.ops_cap;
    x.sc_tau = C.sc_tau + (size_t)chain * C.ops_cap; x.sc_code = C.sc_code + (size_t)chain * C.ops_cap; x.sc_sst = C.sc_sst + (size_t)chain * C.ops_cap;
    x.tc = C.tau_c + (size_t)chain * T.nf * C.kcap; x.ta = C.tau_a + (size_t)chain * T.nf * C.kcap;
    x.sc_tc = C.sc_tc + (size_t)chain * T.nf * C.kcap; x.sc_ta = C.sc_ta + (size_t)chain * T.nf * C.kcap;
    x.Mbase = C.M + (size_t)chain * (T.nf + 1) * ((size_t)C.kcap * C.ld);
    x.ih = C.ih + (size_t)chain * IH_STRIDE; x.dh = C.dh + (size_t)chain * DH_STRIDE;
    x.cnt = C.counters + (size_t)chain * C_NCNT;
    x.wQ = wscratch; x.wR = wscratch + C.kcap; x.wV = wscratch + 2 * C.kcap; x.wI = iscratch;
    if (C.dmax > 1) {
        x.slot = C.op_slot + (size_t)chain * C.ops_cap; x.freel = C.free_slot + (size_t)chain * C.ops_cap;
        x.ket = C.ket + (size_t)chain * C.ops_cap * C.mstride; x.bra = C.bra + (size_t)chain * C.ops_cap * C.mstride;
        x.ketlg = C.ketlg + (size_t)chain * C.ops_cap * C.dmax; x.bralg = C.bralg + (size_t)chain * C.ops_cap * C.dmax;
        x.pt = C.pt + (size_t)chain * 2 * C.dmax;
        x.ldx = C.ldx; x.mX = mscratch; x.mlg = mscratch + C.xdoubles; x.mpt = x.mlg + mat_cols(C.dmax); x.mpw = x.mpt + 2 * mat_cols(C.dmax); x.mY = x.mpw + MAT_DMAX;
        x.kv = x.ih[IH_KV]; x.bv = x.ih[IH_BV]; x.nfree = x.ih[IH_NFREE];
    }
    x.n = x.ih[IH_NOPS]; x.outer = x.ih[IH_OUTER]; x.wpos = x.ih[IH_WPOS]; x.npairs = x.ih[IH_NPAIRS];
    x.perm = x.ih[IH_PERM]; x.palt = x.ih[IH_PALT]; x.err = x.ih[IH_ERR];
    x.ostate = x.ih[IH_OSTATE]; x.only = R.statesampling ? x.ostate : -1;
    x.det_log = x.dh[DH_DETLOG]; x.det_sign = x.dh[DH_DETSIGN]; x.tr_log = x.dh[DH_TRLOG]; x.tr_sign = x.dh[DH_TRSIGN]; x.cfgsign = x.dh[DH_CFGSIGN];
    x.rng.init(R.seed, R.same_key ? 0u : R.chain_base + (unsigned)chain, C.rng_ctr[chain]);
    set_window_edges(T, R, x);
    x.log = nullptr;
}
W2B_HDN void ctx_store(const Chains& C, int chain, Ctx& x) {
    wsync();
    if (lane_id() == 0) {
        x.ih[IH_NOPS] = x.n; x.ih[IH_OUTER] = x.outer; x.ih[IH_WPOS] = x.wpos; x.ih[IH_NPAIRS] = x.npairs;
        x.ih[IH_PERM] = x.perm; x.ih[IH_PALT] = x.palt; x.ih[IH_ERR] = x.err; x.ih[IH_OSTATE] = x.ostate;
        if (C.dmax > 1) { x.ih[IH_KV] = x.kv; x.ih[IH_BV] = x.bv; x.ih[IH_NFREE] = x.nfree; }
        x.dh[DH_DETLOG] = x.det_log; x.dh[DH_DETSIGN] = x.det_sign; x.dh[DH_TRLOG] = x.tr_log; x.dh[DH_TRSIGN] = x.tr_sign; x.dh[DH_CFGSIGN] = x.cfgsign;
        C.rng_ctr[chain] = x.rng.ctr;
    }
    wsync();
}

// ------------------------------------------------------------------ phase scheduler of one CTA (CTQMC.F90:8213-8258, 8428-8533)
// Used for 1x1 quantum-number blocks (C1, C2, C5), where a move is short and a global move (LU of every flavour's matrix)
// costs 100-1000 moves: one persistent kernel, no launch per round.  Blocked traces use the round kernels further down.
// One CTA advances a batch of `nloc` chains by `nsteps` moves each.  Chains are independent Markov chains, so the order
// in which *different* chains execute their moves is free; the scheduler uses that freedom to keep all warps of the CTA in
// the same phase of the algorithm (shared instruction-cache lines, no cross-warp code divergence):
//   round := schedule (every runnable chain draws its move type, reference thresholds CTQMC.F90:8437-8488)
//            -> propose (all insertions, then all removals; warps take chains from a shared-memory work list)
//            -> decide  (determinant ratio, trace, Metropolis, rank-1 update for the proposals that survived)
//   Global moves and tau-shifts (1 % of the moves, 10-100x the cost) park their chain; parked chains are served together
//   in a `rare` phase once a warp-load of them has accumulated (or nothing else is runnable).
// Within a chain the sequence of moves and RNG draws is exactly the reference's.
struct CtaShared {
    int nA, nR, nA2, nR2, nG, cA, cR, cA2, cR2, cG, cA2h, cR2h;
    int* left; unsigned char* parked; unsigned char* kind;
    short* lA; short* lR; short* lA2; short* lR2; short* lG;
    Prop* prop;
};

// Pull the hot records of the chain a warp will work on next into L1 while it is still busy with the current one:
// every lane prefetches one 128-byte line (operator times / superstates / codes, scalar state, counters, the first line
// of each flavour's creator / annihilator times).  For a decide-phase item also the rows of the touched flavour's M.
W2B_HD void prefetch_chain(const Tables& T, const Chains& C, int chain) {
#if W2B_DEVICE
    const int l = lane_id();
    const char* p = nullptr;
    if (l < 8) p = (const char*)(C.op_tau + (size_t)chain * C.ops_cap) + l * 128;
    else if (l < 10) p = (const char*)(C.op_sst + (size_t)chain * C.ops_cap) + (l - 8) * 128;
    else if (l == 10) p = (const char*)(C.op_code + (size_t)chain * C.ops_cap);
    else if (l < 13) p = (const char*)(C.ih + (size_t)chain * IH_STRIDE) + (l - 11) * 128;
    else if (l == 13) p = (const char*)(C.dh + (size_t)chain * DH_STRIDE);
    else if (l == 14) p = (const char*)(C.counters + (size_t)chain * C_NCNT);
    else if (l - 15 < 2 * T.nf) { const int f = (l - 15) >> 1; p = (const char*)(((l - 15) & 1 ? C.tau_a : C.tau_c) + ((size_t)chain * T.nf + f) * C.kcap); }
    if (p) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}
W2B_HD void prefetch_minv(const Tables& T, const Chains& C, int chain, int f) {
#if W2B_DEVICE
    const int* ih = C.ih + (size_t)chain * IH_STRIDE;
    const int k = ih[IH_KF + f];
    const char* m = (const char*)(C.M + ((size_t)chain * (T.nf + 1) + ih[IH_MSLOT + f]) * ((size_t)C.kcap * C.ld));
#pragma unroll 1
    for (int r = lane_id(); r < k; r += W2B_WARP) asm volatile("prefetch.global.L1 [%0];" ::"l"(m + (size_t)r * C.ld * 8));
#endif
}

// bookkeeping after every move: window shift (checked after EVERY move, CTQMC.F90:8255/8531), step counter, quota
W2B_HDN void post_move_cta(const Tables& T, const Run& R, Ctx& x, CtaShared& S, int c) {
    wsync();
    const unsigned long long attempts = x.cnt[C_TRYADD] + x.cnt[C_TRYREM];
    if (!R.freeze_window && attempts % (unsigned long long)R.nslide == 0ull) {         // shift_window Trace.F90:1652-1678
        if (R.num_windows > 1) {
            x.wpos += 1; if (x.wpos == R.num_windows) x.wpos = 0;
            set_window_edges(T, R, x);
            x.npairs = -1;
        }
    }
    if (lane_id() == 0) { if (x.log) x.log[16] = x.wpos; x.cnt[C_STEPS] += 1ull; S.left[c] -= 1; }
    wsync();
}


W2B_HD void cta_sweep(const Tables& T, const Run& R, const Chains& C, const Accum& A, int chain_lo, int nloc, long long nsteps,
                      CtaShared& S, double* wscratch, int* iscratch, double* mscratch, int log_chain, double* logbuf, Ctx& x) {
    const int W = cta_nwarps(), lane = lane_id();
#pragma unroll 1
    for (int c = cta_tid(); c < nloc; c += cta_nthreads()) { S.left[c] = (int)nsteps; S.parked[c] = 0; }
    if (cta_tid() == 0) { S.nG = 0; S.cG = 0; }
    cta_sync();
    const double thr_add = (1.0 - R.pshift - R.pglob) / 2.0, thr_rem = 1.0 - R.pshift - R.pglob, thr_glob = 1.0 - R.pshift;
#pragma unroll 1
    for (;;) {
        if (cta_tid() == 0) { S.nA = 0; S.nA2 = S.nR2 = 0; S.cA = S.cA2 = S.cR2 = S.cA2h = S.cR2h = 0; }
        cta_sync();
        // ---- propose phase: every runnable chain draws move types (reference thresholds CTQMC.F90:8437-8488, plus the two
        //      draws spent on 4-operator / outer moves) and proposes until one proposal survives the quantum-number check
#pragma unroll 1
        for (int c = wgrab(&S.cA), cnext; c < nloc; c = cnext) {
            cnext = wgrab(&S.cA);
            if (S.left[c] <= 0 || S.parked[c]) continue;
            if (R.prefetch && cnext < nloc) prefetch_chain(T, C, chain_lo + cnext);
            const int chain = chain_lo + c;
            ctx_bind(T, R, C, chain, x, wscratch, iscratch, mscratch);
            if (lane == 0) wgrab_one(&S.nA);                              // S.nA counts the chains that were runnable this round
#pragma unroll 1
            for (int run = 0; run < R.propose_run && S.left[c] > 0; ++run) {
                const double r = x.rng.next();
                if (!(r <= thr_rem)) {                                    // global move / tau shift: park for the rare phase
                    if (lane == 0) { S.kind[c] = (unsigned char)(r <= thr_glob ? 3 : 4); S.parked[c] = 1; S.lG[wgrab_one(&S.nG)] = (short)c; }
                    break;
                }
                if (R.pouter > 0.0 || R.p4op > 0.0) {                     // `if (grnd() >= Percentage4OperatorMove) StepAdd(grnd() < PercentageOuterMove) else StepAdd4` (CTQMC.F90:8441-8449)
                    int special = 0;
                    if (!(x.rng.next() >= R.p4op)) special = r < thr_add ? 7 : 8;
                    else if (x.rng.next() < R.pouter) special = r < thr_add ? 5 : 6;
                    if (special) {                                        // 4-operator / outer pair move: parked, served like a global move
                        if (lane == 0) { S.kind[c] = (unsigned char)special; S.parked[c] = 1; S.lG[wgrab_one(&S.nG)] = (short)c; }
                        break;
                    }
                } else x.rng.skip(2);
                if (chain == log_chain) x.log = logbuf + (size_t)(nsteps - S.left[c]) * LOGW;
                bool go;
                if (r < thr_add) { go = add_propose<false>(T, R, C, x, S.prop[c]); if (go && lane == 0) S.lA2[wgrab_one(&S.nA2)] = (short)c; }
                else { go = rem_propose<false>(T, R, C, x, S.prop[c]); if (go && lane == 0) S.lR2[wgrab_one(&S.nR2)] = (short)c; }
                if (go) break;
                post_move_cta(T, R, x, S, c);
            }
            ctx_store(C, chain, x);
        }
        cta_sync();
        const int nrun = S.nA;             // read before thread 0 may reset it at the top of the next round
        bool rare;
        if (nrun == 0) {
            if (S.nG == 0) break;                                         // every chain has used its quota
            rare = true;
        } else {
            // ---- decide phase: determinant ratio, trace, Metropolis, rank-1 update
            const int na2 = S.nA2, nr2 = S.nR2;
            // blocked traces: chains whose outer block is larger than 1x1 cost 10-100x more -> serve them first (longest
            // processing time first keeps the tail of the phase short); pass 0 = heavy, pass 1 = the rest
#pragma unroll 1
            for (int pass = (T.dmax > 1 ? 0 : 1); pass < 2; ++pass) {
                int* curA = pass == 0 ? &S.cA2h : &S.cA2; int* curR = pass == 0 ? &S.cR2h : &S.cR2;
#pragma unroll 1
                for (int i = wgrab(curA); i < na2; i = wgrab(curA)) {
                    const int c = S.lA2[i], chain = chain_lo + c;
                    if (T.dmax > 1 && (ldg(&T.dim[C.ih[(size_t)chain * IH_STRIDE + IH_OUTER]]) > 1) != (pass == 0)) continue;
                    ctx_bind(T, R, C, chain, x, wscratch, iscratch, mscratch);
                    if (chain == log_chain) x.log = logbuf + (size_t)(nsteps - S.left[c]) * LOGW;
                    add_decide<false>(T, R, C, A, x, S.prop[c]);
                    post_move_cta(T, R, x, S, c);
                    ctx_store(C, chain, x);
                }
#pragma unroll 1
                for (int i = wgrab(curR); i < nr2; i = wgrab(curR)) {
                    const int c = S.lR2[i], chain = chain_lo + c;
                    if (T.dmax > 1 && (ldg(&T.dim[C.ih[(size_t)chain * IH_STRIDE + IH_OUTER]]) > 1) != (pass == 0)) continue;
                    ctx_bind(T, R, C, chain, x, wscratch, iscratch, mscratch);
                    if (chain == log_chain) x.log = logbuf + (size_t)(nsteps - S.left[c]) * LOGW;
                    rem_decide<false>(T, R, C, A, x, S.prop[c]);
                    post_move_cta(T, R, x, S, c);
                    ctx_store(C, chain, x);
                }
            }
            cta_sync();
            rare = S.nG >= 2 * W || (S.nG > 0 && nrun < W);          // two warp-loads: the cheap tau shifts fill the tail behind the global moves
        }
        if (rare) {
            // ---- rare phase: global moves, tau shifts and outer pair moves of all parked chains, each whole on one warp.  (Round 2
            // staged the dense linear algebra of large global moves through warp GROUPS eliminating in shared memory; the blocked
            // tensor-pipe routine invert_blocked made one warp per matrix 1.5x faster for the whole sweep at C5 -- 190 -> 128 ms --
            // because 32 matrices are in flight per SM instead of two, and the five stages and their barriers are gone.)
            const int ng = S.nG;
            cta_sync();
            // longest first: global moves (two determinants + two inverses), then everything else; tickets [0, ng) = first pass, [ng, 2 ng) = second
#pragma unroll 1
            for (int t = wgrab(&S.cG); t < 2 * ng; t = wgrab(&S.cG)) {
                const int i = t < ng ? t : t - ng;
                const int c = S.lG[i], chain = chain_lo + c;
                if (((int)S.kind[c] == 3) != (t < ng)) continue;
                ctx_bind(T, R, C, chain, x, wscratch, iscratch, mscratch);
                if (chain == log_chain) x.log = logbuf + (size_t)(nsteps - S.left[c]) * LOGW;
                if ((int)S.kind[c] >= 7) step_four(T, R, C, x, C.rjob[chain], (int)S.kind[c]);
                else if ((int)S.kind[c] >= 5) step_outer(T, R, C, A, x, S.prop[c], C.rjob[chain], (int)S.kind[c]);
                else step_rare(T, R, C, x, C.rjob[chain], (int)S.kind[c]);
                post_move_cta(T, R, x, S, c);
                ctx_store(C, chain, x);
                if (lane == 0) S.parked[c] = 0;
            }
            cta_sync();
            if (cta_tid() == 0) { S.nG = 0; S.cG = 0; }
            cta_sync();
        }
    }
}


// (L2 prefetch bursts of a chain's records at the start of a round item were measured: no effect, 1 % slower.)
// bookkeeping after every move: window shift (checked after EVERY move, CTQMC.F90:8255/8531), step counter, quota
W2B_HDN void post_move(const Tables& T, const Run& R, const Chains& C, Ctx& x, int chain) {
    wsync();
    const unsigned long long attempts = x.cnt[C_TRYADD] + x.cnt[C_TRYREM];
    if (!R.freeze_window && attempts % (unsigned long long)R.nslide == 0ull) {         // shift_window Trace.F90:1652-1678
        if (R.num_windows > 1) {
            x.wpos += 1; if (x.wpos == R.num_windows) x.wpos = 0;
            set_window_edges(T, R, x);
            x.npairs = -1;
        }
    }
    if (lane_id() == 0) { if (x.log) x.log[16] = x.wpos; x.cnt[C_STEPS] += 1ull; C.left[chain] -= 1; }
    wsync();
}

// ------------------------------------------------------------------ autonomous sweep (1x1 blocks, large orders)
// One warp takes one chain through ALL its moves: context bound once, no CTA phases, no barriers.  The phase scheduler above exists
// because short moves are bound by instruction delivery (every warp at its own place in 10 k instructions of code); at large
// expansion orders (C5, k = 74) a move is a handful of long loops over a 44 kB matrix, the code between them matters little, and
// what the phases cost -- every chain's records re-read per phase, the whole CTA waiting at three barriers per round behind its
// longest item, global moves (5 ms on one warp) served in CTA-wide rare phases -- is most of the step (profiles/r2q_*: 40 % of the
// warp samples at the CTA barrier, and the sweep time did not move when 14 % of its instructions were removed).
W2B_HD void chain_sweep(const Tables& T, const Run& R, const Chains& C, const Accum& A, int chain, long long nsteps, Prop& P,
                        double* wscratch, int* iscratch, double* mscratch, int log_chain, double* logbuf, Ctx& x) {
    const double thr_add = (1.0 - R.pshift - R.pglob) / 2.0, thr_rem = 1.0 - R.pshift - R.pglob, thr_glob = 1.0 - R.pshift;
    ctx_bind(T, R, C, chain, x, wscratch, iscratch, mscratch);
#if defined(W2B_TIMING) && W2B_DEVICE
#define W2B_TICK(slot) { const long long t1_ = clock64(); if (lane_id() == 0) atomicAdd((unsigned long long*)&C.sched[16 + 2 * (slot)], (unsigned long long)(t1_ - t0_)); t0_ = t1_; }
    long long t0_ = clock64();
#else
#define W2B_TICK(slot)
#endif
#pragma unroll 1
    for (long long s = 0; s < nsteps; ++s) {
        if (chain == log_chain) x.log = logbuf + (size_t)s * LOGW;
        const double r = x.rng.next();                                // move type: reference thresholds CTQMC.F90:8437-8488
        if (!(r <= thr_rem)) { const int kind = r <= thr_glob ? 3 : 4; step_rare(T, R, C, x, C.rjob[chain], kind); W2B_TICK(kind == 3 ? 4 : 5) }
        else {
            bool outer = false, four = false;
            if (R.pouter > 0.0 || R.p4op > 0.0) { four = !(x.rng.next() >= R.p4op); if (!four) outer = x.rng.next() < R.pouter; } else x.rng.skip(2);
            if (four) step_four(T, R, C, x, C.rjob[chain], r < thr_add ? 7 : 8);
            else if (outer) step_outer(T, R, C, A, x, P, C.rjob[chain], r < thr_add ? 5 : 6);
            else if (r < thr_add) { const bool go = add_propose<false>(T, R, C, x, P); W2B_TICK(0) if (go) { add_decide<false>(T, R, C, A, x, P); W2B_TICK(1) } }
            else { const bool go = rem_propose<false>(T, R, C, x, P); W2B_TICK(2) if (go) { rem_decide<false>(T, R, C, A, x, P); W2B_TICK(3) } }
        }
        post_move(T, R, C, x, chain);
        W2B_TICK(6)
    }
    ctx_store(C, chain, x);
}

// ------------------------------------------------------------------ round scheduler (CTQMC.F90:8213-8258, 8428-8533)
// Chains are independent Markov chains, so the order in which *different* chains execute their moves is free.  A sweep of
// `nsteps` moves per chain runs as rounds of two kernels over the whole population:
//   k_propose : every runnable chain draws move types (reference thresholds CTQMC.F90:8437-8488) and proposes until one
//               proposal survives the quantum-number check (at most R.propose_run moves); survivors are queued for the
//               decide kernel, global moves / tau shifts for its rare queue.  No shared memory, full occupancy.
//   k_decide  : determinant ratio, blocked trace on the tensor pipe, Metropolis, rank-1 update.  Warps take tickets from
//               one global counter, longest jobs first: rare moves, then outer blocks >= 16, > 1, == 1.
// Each kernel is small enough for the instruction cache and sized for its own register / shared-memory needs.
// Within a chain the sequence of moves and RNG draws is exactly the reference's.
// what one chain's propose phase hands to the round's queues; collected per CTA and flushed with one atomic per queue
// (tens of thousands of chains pushing one by one serialise on the counters' L2 atomics)
struct ProposeOut { int run, left, dq, entry, tq, tentry, ntl, rfin; };      // dq / tq: decide / trace queue (0..2, -1 = none)

// `x` is the warp's chain context: the kernels keep ONE copy per warp in shared memory (every lane holds the same values;
// a per-thread copy lives in local memory, 32x replicated, and its write-back was 1.1 GB of DRAM traffic per propose launch)
W2B_HD void phase_propose(const Tables& T, const Run& R, const Chains& C, int chain, long long nsteps, int parity, int log_chain, double* logbuf,
                          ProposeOut* out, Ctx& x) {
    const int lane = lane_id();
    int left = C.left[chain];
    if (left <= 0) return;
    if (lane == 0) { out->run = 1; out->left = left; }
    const int state = vload_u8(&C.parked[chain]);
    if (state != ST_RUN) {                                        // global move / tau shift in flight on the rare stream
        if (lane == 0) {
            // a move completed on the rare stream re-enters with the NEXT round: its records were written while this kernel was
            // already running and may sit stale in this SM's L1 (lines shared with neighbouring chains)
            if (state == ST_RARE_DONE) C.parked[chain] = ST_RUN;
            if (state == ST_RARE_TRACE) {                         // first half done: this round's trace kernel runs its jobs, the decide kernel the second half
                const int new_outer = *(const volatile int*)&C.rjob[chain].new_outer;
                out->tq = 0; out->tentry = chain | Q_RARE; out->ntl = (ldg(&T.dim[new_outer]) + TRACE_COLS - 1) / TRACE_COLS; out->rfin = 1;
                if (R.statesampling) {                            // only the column tile that holds the proposed outer eigenstate
                    const int ns = *(const volatile int*)&C.rjob[chain].new_state;
                    out->tentry |= ((ns < 0 ? 0 : ns) / TRACE_COLS) << Q_NTSHIFT; out->ntl = 1;
                }
                C.parked[chain] = ST_RARE_QUEUED;
            }
        }
        return;
    }
    const double thr_add = (1.0 - R.pshift - R.pglob) / 2.0, thr_rem = 1.0 - R.pshift - R.pglob, thr_glob = 1.0 - R.pshift;
    ctx_bind(T, R, C, chain, x, nullptr, nullptr, nullptr);
#pragma unroll 1
    for (int run = 0; run < R.propose_run && left > 0; ++run) {
        const double r = x.rng.next();
        if (!(r <= thr_rem)) {                                    // global move / tau shift: rare stream
            if (lane == 0) { C.kind[chain] = (unsigned char)(r <= thr_glob ? 3 : 4); C.parked[chain] = ST_RARE_NEW; }
            break;
        }
        if (R.pouter > 0.0 || R.p4op > 0.0) {                     // `if (grnd() >= Percentage4OperatorMove) StepAdd(grnd() < PercentageOuterMove) else StepAdd4` (CTQMC.F90:8441-8449)
            int special = 0;
            if (!(x.rng.next() >= R.p4op)) special = r < thr_add ? 7 : 8;
            else if (x.rng.next() < R.pouter) special = r < thr_add ? 5 : 6;
            if (special) {                                        // 4-operator / outer pair move: rare stream
                if (lane == 0) { C.kind[chain] = (unsigned char)special; C.parked[chain] = ST_RARE_NEW; }
                break;
            }
        } else x.rng.skip(2);
        if (chain == log_chain) x.log = logbuf + (size_t)(nsteps - left) * LOGW;
        bool go; int entry = chain;
        if (r < thr_add) go = add_propose<false>(T, R, C, x, C.prop[chain]);
        else { go = rem_propose<false>(T, R, C, x, C.prop[chain]); entry |= Q_REM; }
        if (go) {
            if (lane == 0) {
                const int m = T.dmax > 1 ? ldg(&T.dim[x.outer]) : 1;
                out->dq = (entry & Q_REM) ? 1 : 0; out->entry = entry;   // insertions and removals in separate queues: neighbouring warps of the decide kernel run the same code
                if (T.dmax > 1) {                                 // trace jobs: one per 8 outer states, queued by number of steps
                    const Prop& P = C.prop[chain];
                    const int rem = (entry & Q_REM) ? 1 : 0, s = P.i1, e = P.i2 + rem;
                    const int steps = (x.bv > e ? x.bv - e : 0) + (x.kv <= s - 1 ? s - x.kv : 0) + (P.i2 - P.i1 - rem) + (rem ? 0 : 2);
                    out->tq = steps >= 24 ? 0 : (steps >= 8 ? 1 : 2); out->tentry = entry; out->ntl = (m + TRACE_COLS - 1) / TRACE_COLS;
                    if (R.statesampling) { out->tentry |= (x.ostate / TRACE_COLS) << Q_NTSHIFT; out->ntl = 1; }       // the sampled state's column tile
                }
            }
            break;
        }
        post_move(T, R, C, x, chain);
        --left;
    }
    ctx_store(C, chain, x);
}
// one warp appends the `count` results of a CTA to the round's queues
W2B_HD void propose_flush(const Chains& C, int parity, const ProposeOut* outs, int count) {
    int* sc = C.sched + parity * SC_STRIDE;
    const int lane = lane_id();
#pragma unroll 1
    for (int base = 0; base < count; base += W2B_WARP) {
        const int i = base + lane;
        ProposeOut o = {0, 0, -1, 0, -1, 0, 0, 0};
        if (i < count) o = outs[i];
        const int nrun = wsum(o.run), ml = wmax(o.run ? o.left : 0);
        if (lane == 0 && nrun) { atomic_add_int(&sc[SC_NRUN], nrun); atomic_max_int(&sc[SC_MAXLEFT], ml); }
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
            {   // decide queue c
                const int mine = o.dq == c ? 1 : 0, incl = wscan_incl(mine), tot = wbcast(incl, W2B_WARP - 1);
                if (tot) {
                    int b = 0; if (lane == 0) b = atomic_add_int(&sc[SC_NQ0 + c], tot);
                    b = wbcast(b, 0);
                    if (mine) (c == 0 ? C.q_dec0 : (c == 1 ? C.q_dec1 : C.q_dec2))[b + incl - 1] = o.entry;
                }
            }
            {   // trace queue c: ntl consecutive entries per chain
                const int cnt = o.tq == c ? o.ntl : 0, incl = wscan_incl(cnt), tot = wbcast(incl, W2B_WARP - 1);
                if (tot) {
                    int b = 0; if (lane == 0) b = atomic_add_int(&sc[SC_NT0 + c], tot);
                    b = wbcast(b, 0);
                    int* q = (c == 0 ? C.q_tr0 : (c == 1 ? C.q_tr1 : C.q_tr2)) + (size_t)parity * 4 * C.nchains;
#pragma unroll 1
                    for (int t = 0; t < cnt; ++t) q[b + incl - cnt + t] = o.tentry | (t << Q_NTSHIFT);
                }
            }
        }
        const int mine = o.rfin ? 1 : 0, incl = wscan_incl(mine), tot = wbcast(incl, W2B_WARP - 1);
        if (tot) {
            int b = 0; if (lane == 0) b = atomic_add_int(&sc[SC_NRFIN], tot);
            b = wbcast(b, 0);
            if (mine) C.q_rfin[(size_t)parity * C.nchains + b + incl - 1] = o.tentry & Q_IDMASK;
        }
    }
}

// one ticket of the trace kernel (blocked traces only): tickets run over the three trace queues, longest jobs first
W2B_HD void phase_trace(const Tables& T, const Chains& C, int ticket, int parity, double* mscratch) {
    const int* sc = C.sched + parity * SC_STRIDE;
    const size_t qo = (size_t)parity * 4 * C.nchains;
    const int n0 = sc[SC_NT0], n1 = sc[SC_NT1];
    const int entry = ticket < n0 ? C.q_tr0[qo + ticket] : (ticket < n0 + n1 ? C.q_tr1[qo + ticket - n0] : C.q_tr2[qo + ticket - n0 - n1]);
    trace_job(T, C, entry, mscratch);
}

// Global moves / tau shifts (1 % of the moves, 10-1000x the cost of a pair move) run on their own stream beside the round
// pipeline: the rare kernel scans the chain states, takes the chains a propose kernel left in ST_RARE_NEW and runs
//   1x1 blocks:     the whole move; the chain returns to ST_RUN;
//   blocked traces: the first half (everything up to the local trace), then ST_RARE_TRACE: the next propose kernel queues the
//                   trace jobs and the second half for its round (trace kernel / decide kernel).
// The round pipeline never waits for this stream; a chain whose rare move is in flight simply sits out a few rounds.
W2B_HD void phase_rare(const Tables& T, const Run& R, const Chains& C, const Accum& A, int chain, long long nsteps,
                       int log_chain, double* logbuf, double* wscratch, int* iscratch, double* mscratch, Ctx& x) {
    if (vload_u8(&C.parked[chain]) != ST_RARE_NEW) return;
    ctx_bind(T, R, C, chain, x, wscratch, iscratch, mscratch);
    if (chain == log_chain) x.log = logbuf + (size_t)(nsteps - C.left[chain]) * LOGW;
    RareJob& J = C.rjob[chain];
    const int kind = C.kind[chain];
    bool pending = false;
    if (kind >= 7) {                                                 // 4-operator moves (Percentage4OperatorMove)
        if (T.dmax > 1) pending = four_prepare(T, R, C, x, J, kind);
        else step_four(T, R, C, x, J, kind);
    } else if (kind >= 5) {                                          // outer pair moves (PercentageOuterMove)
        if (T.dmax > 1) pending = outer_prepare(T, R, C, x, C.prop[chain], J, kind);
        else step_outer(T, R, C, A, x, C.prop[chain], J, kind);
    } else if (T.dmax > 1) {
        pending = kind == 3 ? step_glob_prepare(T, R, C, x, J) : step_shift_prepare(T, R, C, x, J);
        if (pending && kind == 4 && lane_id() == 0) J.stored = 1;    // tau shifts are accepted with probability ~1: cache the kets of the new sequence
    } else step_rare(T, R, C, x, J, kind);
    if (!pending) post_move(T, R, C, x, chain);
    ctx_store(C, chain, x);
    fence_device();
    if (lane_id() == 0) vstore_u8(&C.parked[chain], (unsigned char)(pending ? ST_RARE_TRACE : ST_RARE_DONE));
}

// one ticket of the decide kernel: tickets [0, nrare) are rare moves, then the three decide queues
W2B_HD void phase_decide(const Tables& T, const Run& R, const Chains& C, const Accum& A, int ticket, int parity, long long nsteps,
                         int log_chain, double* logbuf, double* wscratch, int* iscratch, double* mscratch, Ctx& x) {
    const int lane = lane_id();
    const int* sc = C.sched + parity * SC_STRIDE;
    const int nrare = sc[SC_NRFIN], n0 = sc[SC_NQ0], n1 = sc[SC_NQ1];
    int entry; bool rare = false;
    if (ticket < nrare) { entry = C.q_rfin[(size_t)parity * C.nchains + ticket]; rare = true; }
    else if (ticket < nrare + n0) entry = C.q_dec0[ticket - nrare];
    else if (ticket < nrare + n0 + n1) entry = C.q_dec1[ticket - nrare - n0];
    else entry = C.q_dec2[ticket - nrare - n0 - n1];
    const int chain = entry & Q_IDMASK;
    ctx_bind(T, R, C, chain, x, wscratch, iscratch, mscratch);
    if (chain == log_chain) x.log = logbuf + (size_t)(nsteps - C.left[chain]) * LOGW;
    if (T.dmax > 1) x.mpt = C.ptc + (size_t)chain * 2 * C.dmax;      // parts of the proposed configuration, left by the trace kernel
    if (rare) {                                                   // second half of a global move / tau shift: the trace kernel has left the parts in C.ptc
        RareJob& J = C.rjob[chain];
        double tn_log, tn_sign;
        x.mpt = C.ptc + (size_t)chain * 2 * C.dmax;
        parts_sum(x.mpt, ldg(&T.dim[J.new_outer]), tn_log, tn_sign, R.statesampling ? J.new_state : -1);
        const int kind = C.kind[chain];
        if (kind == 3) step_glob_finish(T, R, C, x, J, tn_log, tn_sign);
        else if (kind == 4) step_shift_finish(T, R, C, x, J, tn_log, tn_sign);
        else if (kind == 5) add_decide<true>(T, R, C, A, x, C.prop[chain]);
        else if (kind == 6) rem_decide<true>(T, R, C, A, x, C.prop[chain]);
        else four_finish(T, R, C, x, J, tn_log, tn_sign);
    }
    else if (entry & Q_REM) rem_decide<false>(T, R, C, A, x, C.prop[chain]);
    else add_decide<false>(T, R, C, A, x, C.prop[chain]);
    post_move(T, R, C, x, chain);
    ctx_store(C, chain, x);
    if (rare && lane == 0) C.parked[chain] = ST_RUN;
}

// ------------------------------------------------------------------ autocorrelation accumulator of G(tau) (AccumGtau = 2)
// Restates autocorracc_add (Accumulators.F90:285-330; logarithmic binning with Welford updates, arXiv:1810.05079 Sec. III.D)
// for one component of one chain-group bin.  The series element of measurement n is the bin's accumulator increment divided
// by its number of chains (the group mean of what the reference's single chain puts into `buffer`).
// dat: [level][3][ncomp] of this bin (spill, mean, sqdiff).
W2B_HD void autocorr_add_one(double val, long long n, int nlevels, double* dat, int ncomp, int comp) {
    int lvl = 0;
    double* sp = dat + comp; double* mn = dat + ncomp + comp; double* sq = dat + 2 * ncomp + comp;
    {
        *sp += val;
        const double incr1 = val - *mn; *mn += incr1 / (double)(n + 1);
        const double incr2 = val - *mn; *sq += incr1 * incr2;
    }
    while (n & 1) {                      // the lowest bits are all ones: roll the finished bin over to the next level
        if (lvl == nlevels - 1) break;
        n /= 2; ++lvl;
        const double v = *sp / 2.0; *sp = 0.0;
        sp += (size_t)3 * ncomp; mn += (size_t)3 * ncomp; sq += (size_t)3 * ncomp;
        *sp += v;
        const double incr1 = v - *mn; *mn += incr1 / (double)(n + 1);
        const double incr2 = v - *mn; *sq += incr1 * incr2;
    }
}

// init_Trace (Trace.F90:340-701): empty configuration, outer superstate drawn with the empty-trace weights
W2B_HD void init_chain(const Tables& T, const Run& R, const Chains& C, int chain) {
    if (lane_id() != 0) return;
    int* ih = C.ih + (size_t)chain * IH_STRIDE; double* dh = C.dh + (size_t)chain * DH_STRIDE;
    // the reference seeds once (CtHybSolver.__init__ -> init_rnd -> sgrnd) and lets the generator run on across init_solver calls:
    // solve number `stream` since init_rnd starts at counter stream << 44 of the chain's stream (0 for the first solve)
    Philox rng; rng.init(R.seed, R.same_key ? 0u : R.chain_base + (unsigned)chain, (unsigned long long)R.stream << STREAM_SHIFT);
    const double rand = rng.next();
    double acc = 0.0; int pick = T.nsst - 1;
#pragma unroll 1
    for (int i = 0; i < T.nsst && !R.statesampling; ++i) { double w = T.empty_w[i]; acc += w; pick = i; if (acc >= rand && w != 0.0) break; }
    int pstate = 0;
    if (R.statesampling) {                 // initial outer (superstate, eigenstate) ~ empty-trace value of the state (Trace.F90:612-643)
        bool done = false;
#pragma unroll 1
        for (int i = 0; i < T.nsst && !done; ++i) {
#pragma unroll 1
            for (int j = 0; j < T.dim[i]; ++j) {
                const double w = w_exp(-(T.evals[T.offset[i] + j] - T.egs) * T.beta) / T.empty_norm;
                acc += w; pick = i; pstate = j;
                if (acc >= rand && w != 0.0) { done = true; break; }
            }
        }
    }
#pragma unroll 1
    for (int i = 0; i < IH_STRIDE; ++i) ih[i] = 0;
    ih[IH_OUTER] = pick; ih[IH_NPAIRS] = 0; ih[IH_OSTATE] = pstate;
#pragma unroll 1
    for (int f = 0; f <= T.nf; ++f) ih[IH_MSLOT + f] = f;
    dh[DH_DETLOG] = 0.0; dh[DH_DETSIGN] = 1.0; dh[DH_TRLOG] = R.statesampling ? -(T.evals[T.offset[pick] + pstate] - T.egs) * T.beta : T.empty_log[pick]; dh[DH_TRSIGN] = 1.0; dh[DH_CFGSIGN] = -1.0;
    C.rng_ctr[chain] = rng.ctr;
#pragma unroll 1
    for (int i = 0; i < C_NCNT; ++i) C.counters[(size_t)chain * C_NCNT + i] = 0ull;
    if (C.dmax > 1) {
        ih[IH_KV] = 0; ih[IH_BV] = 0; ih[IH_NFREE] = C.ops_cap;
#pragma unroll 1
        for (int i = 0; i < C.ops_cap; ++i) { C.op_slot[(size_t)chain * C.ops_cap + i] = (unsigned short)i; C.free_slot[(size_t)chain * C.ops_cap + i] = (unsigned short)(C.ops_cap - 1 - i); }
        const double* ev = T.evals + T.offset[pick];
#pragma unroll 1
        for (int j = 0; j < T.dim[pick]; ++j) { C.pt[(size_t)chain * 2 * C.dmax + 2 * j] = -(ev[j] - T.egs) * T.beta; C.pt[(size_t)chain * 2 * C.dmax + 2 * j + 1] = 1.0; }
    }
}

// ------------------------------------------------------------------ measurement (CTQMC.F90:8541-8581)
constexpr int LEG_CHUNK = 25;
constexpr int LEG_CHUNK_BIG = 50;
// blocked-trace part of one measurement: per-state histograms (MeasOuterHistograms CTQMC.F90:3285-3339) and the density
// matrix at beta/2 (update_b2_states_eb Trace.F90:6213-6469 + MeasDensityMatrix_beta_half CTQMC.F90:3088-3151)
W2B_HDN void measure_blocked(const Tables& T, const Run& R, const Chains& C, const Accum& A, Ctx& x, double* acc) {
    const int m = ldg(&T.dim[x.outer]), off = ldg(&T.offset[x.outer]);
    double csum = 0.0;
    const int only = x.only;                // state sampling: the one outer eigenstate of the configuration (CTQMC.F90:3294-3339, 3106-3110)
#pragma unroll 1
    for (int j = lane_id(); j < m; j += W2B_WARP) {
        if (only >= 0 && j != only) continue;
        const double e = x.pt[2 * j + 1] != 0.0 ? w_exp(x.pt[2 * j]) : 0.0;
        atomic_add(&acc[A.o_outerst + off + j], 1.0);
        atomic_add(&acc[A.o_signst + off + j], x.tr_sign * x.pt[2 * j + 1] * x.cfgsign);
        atomic_add(&acc[A.o_contribst + off + j], e);
        csum += e;
    }
    csum = wsum(csum);
    if (lane_id() == 0) atomic_add(&acc[A.o_contribsst + x.outer], csum);
    if (!R.meas_dm) return;
    if (x.n == 0) {
        const int doff = ldg(&T.dens_off[x.outer]);
#pragma unroll 1
        for (int j = lane_id(); j < m; j += W2B_WARP) if (only < 0 || j == only)
            atomic_add(&acc[A.o_dens + doff + j + j * m], -x.cfgsign * x.pt[2 * j + 1] * x.tr_sign * w_exp(x.pt[2 * j] - x.tr_log));
        return;
    }
    const double bh = T.beta / 2.0;
    const int ih = count_le(x.tau, x.n, bh);                 // operators with tau <= beta/2 belong to the ket
    if (C.meas_pre) { if (x.kv < ih) x.kv = ih; if (x.bv > ih) x.bv = ih; }     // extended by the trace jobs of Engine::measure
    extend_kets(T, C, x, ih - 1);
    extend_bras(T, C, x, ih);
    const int sst = ih == 0 ? x.outer : (int)x.sst[ih - 1];
    const int d = ldg(&T.dim[sst]);
    const double* ev = T.evals + ldg(&T.offset[sst]);
    const double e0 = ldg(&ev[0]);
    const double tauR = ih == 0 ? 0.0 : x.tau[ih - 1], tauL = ih == x.n ? T.beta : x.tau[ih];
    double* K = x.mX; double* B = x.mY; double* klg = x.mlg; double* blg = x.mpt; double* fac = x.mpt + C.dmax;
    if (ih == 0) mat_identity_compact(K, klg, m);
    else { mat_copy(K, ket_slot(C, x, ih - 1), d * m); mat_copy(klg, x.ketlg + (size_t)x.slot[ih - 1] * C.dmax, m); }
    if (ih == x.n) mat_identity_compact(B, blg, m);
    else { mat_copy(B, bra_slot(C, x, ih), d * m); mat_copy(blg, x.bralg + (size_t)x.slot[ih] * C.dmax, m); }
    wsync();
#pragma unroll 1
    for (int e = lane_id(); e < d * m; e += W2B_WARP) {
        const int c = e % d; const double de = ldg(&ev[c]) - e0;
        K[e] *= w_exp(-de * (bh - tauR)); B[e] *= w_exp(-de * (tauL - bh));
    }
#pragma unroll 1
    for (int j = lane_id(); j < m; j += W2B_WARP) {
        const double kl = klg[j] - (e0 - T.egs) * (bh - tauR), bl = blg[j] - (e0 - T.egs) * (tauL - bh);
        fac[j] = (only >= 0 && j != only) ? 0.0 : x.tr_sign * (-x.cfgsign) * w_exp(kl + bl - x.tr_log);      // 1 / trval(Trace / TLogTr(kl + bl, -cfgsign))
    }
    wsync();
    const int doff = ldg(&T.dens_off[sst]);
#pragma unroll 1
    for (int e = lane_id(); e < d * d; e += W2B_WARP) {
        const int a = e % d, c = e / d;
        double v = 0.0;
#pragma unroll 1
        for (int j = (only >= 0 ? only : 0); j < (only >= 0 ? only + 1 : m); ++j) v += B[a + (size_t)j * d] * K[c + (size_t)j * d] * fac[j];
        atomic_add(&acc[A.o_dens + doff + a + c * d], v);
    }
    wsync();
}

// gleg_sh (optional): the warp's private shared-memory accumulator of the Legendre coefficients [nf][nlegmax] (all chains of a
// CTA belong to one bin); the CTA flushes its warps' accumulators to the bin's slab once, instead of nf * NLegMax global atomics per chain
W2B_HD void measure_chain(const Tables& T, const Run& R, const Chains& C, const Accum& A, int chain, double* wscratch, int* iscratch, double* mscratch,
                          double* gleg_sh = nullptr) {
    const int lane = lane_id();
    const int* ih = C.ih + (size_t)chain * IH_STRIDE; const double* dh = C.dh + (size_t)chain * DH_STRIDE;
    const int n = ih[IH_NOPS], outer = ih[IH_OUTER];
    const double cfgsign = dh[DH_CFGSIGN];
    double* acc = A.base + (size_t)((R.chain_base + (unsigned)chain) % (unsigned)R.nbins) * A.stride;
    const double* tau = C.op_tau + (size_t)chain * C.ops_cap; const short* sst = C.op_sst + (size_t)chain * C.ops_cap;
    if (lane == 0) {
        atomic_add(&acc[A.o_sign], cfgsign); atomic_add(&acc[A.o_cnt], 1.0);                 // MeasSign 3343
        atomic_add(&acc[A.o_outer + outer], 1.0); atomic_add(&acc[A.o_signsst + outer], cfgsign);   // MeasOuterHistograms 3285-3339
        if (T.dmax == 1) {
            atomic_add(&acc[A.o_contribsst + outer], w_exp(dh[DH_TRLOG]));
            int off = T.offset[outer];
            atomic_add(&acc[A.o_outerst + off], 1.0); atomic_add(&acc[A.o_signst + off], cfgsign);  // Trace.sign*parttrace.sign = 1 for a 1x1 block
            atomic_add(&acc[A.o_contribst + off], w_exp(dh[DH_TRLOG]));
        }
#pragma unroll 1
        for (int f = 0; f < T.nf; ++f) {                                                       // MeasExpOrder 3264-3282
            int k = ih[IH_KF + f]; if (k > R.maxhisto) k = R.maxhisto;
            atomic_add(&acc[A.o_histo + f * (R.maxhisto + 1) + k], 1.0);
        }
    }
    if (T.dmax > 1) {
        Ctx x;
        ctx_bind(T, R, C, chain, x, wscratch, iscratch, mscratch);
        wsync();
        measure_blocked(T, R, C, A, x, acc);
        ctx_store(C, chain, x);
    } else if (R.meas_dm) {                                                                   // MeasDensityMatrix_beta_half 3088-3151 (1x1 blocks)
        int ihalf = count_le(tau, n, T.beta / 2.0);
        if (lane == 0) {
            int s = ihalf == 0 ? outer : (int)sst[ihalf - 1];
            atomic_add(&acc[A.o_dens + T.offset[s]], -cfgsign);
        }
    }
    // MeasGtauRem 2252-2440: all (annihilator, creator) pairs of every flavour
    const double beta = T.beta;
    if (gleg_sh) {
        // GPU mapping: one lane per (flavour, creator, annihilator) pair over ALL flavours at once (a 5-orbital chain has ~7
        // pairs per flavour: a lane-per-pair loop inside one flavour leaves 3/4 of the warp idle); each lane runs the whole
        // three-term recurrence (CTQMC.F90:2407-2420, coefficients tabulated) and adds into the CTA's shared accumulator --
        // warp-reduced first when the whole warp works on one flavour (large orders)
        // flavours with more than one warp-load of pairs: all lanes work on the same flavour anyway -- one pair per lane, 50 partial
        // sums per lane in registers over all pairs, one warp reduction per l at the end (two passes over the pairs, the second
        // restarts the recurrence without accumulating).  The recurrence runs on MONIC Legendre polynomials,
        //     Q_l = x Q_{l-1} - g_l Q_{l-2},  g_l = (l-1)^2 / ((2l-1)(2l-3)),  P_l = c_l Q_l,
        // two FP64 operations per step instead of three (CTQMC.F90:2407-2420 divides by l in every step); the leading coefficients
        // c_l multiply the reduced sums once.  At beta = 200 (5600 pairs per flavour) this loop was a third of the whole step.
        for (int f = 0; f < T.nf; ++f) {
            const int k = ih[IH_KF + f];
            if (k * k <= W2B_WARP) continue;
            const double* tc = C.tau_c + ((size_t)chain * T.nf + f) * C.kcap; const double* ta = C.tau_a + ((size_t)chain * T.nf + f) * C.kcap;
            const double* M = C.M + ((size_t)chain * (T.nf + 1) + ih[IH_MSLOT + f]) * ((size_t)C.kcap * C.ld);
            double* g = gleg_sh + f * R.nlegmax;
            const double* gam = T.leg_g;
#pragma unroll 1
            for (int lbase = 0; lbase < R.nlegmax; lbase += LEG_CHUNK_BIG) {
                double part[LEG_CHUNK_BIG];
#pragma unroll
                for (int q = 0; q < LEG_CHUNK_BIG; ++q) part[q] = 0.0;
                int ic = 0, ia = lane;                                  // pair p = ic * k + ia, advanced without divisions
                while (ia >= k) { ia -= k; ++ic; }
                // TWO pairs per lane and iteration (p and p + 32): the recurrence is one dependent DMUL -> DFMA chain per pair and the
                // kernel runs 2 warps per scheduler (238 registers), so a second independent chain per lane doubles what the FP64
                // pipe has in flight
#pragma unroll 1
                for (int p = lane; p < k * k; p += 2 * W2B_WARP) {
                    double MvA = M[(size_t)ic * C.ld + ia], xxA, MvB = 0.0, xxB = 0.0;
                    {
                        const double dtau = ta[ia] - tc[ic]; int idt;
                        if (dtau > 0.0) { idt = (int)(dtau / beta * (double)R.ntau) + 1; xxA = 2.0 / beta * dtau - 1.0; }
                        else { idt = (int)((beta + dtau) / beta * (double)R.ntau) + 1; xxA = 2.0 / beta * (beta + dtau) - 1.0; MvA = -MvA; }
                        MvA *= cfgsign;
                        if (lbase == 0) { if (idt > R.ntau) idt = R.ntau; atomic_add(&acc[A.o_gtau + f * R.ntau + idt - 1], MvA); }
                    }
                    ia += W2B_WARP; while (ia >= k) { ia -= k; ++ic; }
                    if (p + W2B_WARP < k * k) {
                        MvB = M[(size_t)ic * C.ld + ia];
                        const double dtau = ta[ia] - tc[ic]; int idt;
                        if (dtau > 0.0) { idt = (int)(dtau / beta * (double)R.ntau) + 1; xxB = 2.0 / beta * dtau - 1.0; }
                        else { idt = (int)((beta + dtau) / beta * (double)R.ntau) + 1; xxB = 2.0 / beta * (beta + dtau) - 1.0; MvB = -MvB; }
                        MvB *= cfgsign;
                        if (lbase == 0) { if (idt > R.ntau) idt = R.ntau; atomic_add(&acc[A.o_gtau + f * R.ntau + idt - 1], MvB); }
                    }
                    double qm1A = 1.0, q0A = xxA, qm1B = 1.0, q0B = xxB;   // Q_0, Q_1
                    int l0 = 2;
                    if (lbase == 0) { part[0] += MvA + MvB; part[1] += MvA * xxA + MvB * xxB; }
                    else {
#pragma unroll 4
                        for (int l = 2; l < lbase; ++l) {
                            const double gl = -ldg(&gam[l]);
                            const double qnA = fma(xxA, q0A, gl * qm1A), qnB = fma(xxB, q0B, gl * qm1B);
                            qm1A = q0A; q0A = qnA; qm1B = q0B; q0B = qnB;
                        }
                        l0 = 0;
                    }
#pragma unroll
                    for (int q = 0; q < LEG_CHUNK_BIG; ++q) {
                        if (q >= l0 && lbase + q < R.nlegmax) {
                            const double gl = -ldg(&gam[lbase + q]);
                            const double qnA = fma(xxA, q0A, gl * qm1A), qnB = fma(xxB, q0B, gl * qm1B);
                            part[q] = fma(MvB, qnB, fma(MvA, qnA, part[q]));
                            qm1A = q0A; q0A = qnA; qm1B = q0B; q0B = qnB;
                        }
                    }
                    ia += W2B_WARP; while (ia >= k) { ia -= k; ++ic; }
                }
#pragma unroll
                for (int q = 0; q < LEG_CHUNK_BIG; ++q) {
                    const double sum = wsum(part[q]);
                    if (lane == 0 && lbase + q < R.nlegmax) g[lbase + q] += sum * ldg(&T.leg_c[lbase + q]);
                }
            }
            wsync();
        }
        // small flavours: one lane per pair across all of them
        int offs[NF_MAX + 1]; offs[0] = 0;
#pragma unroll 1
        for (int f = 0; f < T.nf; ++f) { const int k2 = ih[IH_KF + f] * ih[IH_KF + f]; offs[f + 1] = offs[f] + (k2 <= W2B_WARP ? k2 : 0); }
        const int total = offs[T.nf];
#pragma unroll 1
        for (int base = 0; base < total; base += W2B_WARP) {
            const int p = base + lane; const bool act = p < total;
            int f = 0;
            if (act) while (p >= offs[f + 1]) ++f;
            double Mv = 0.0, xx = 0.0;
            if (act) {
                const int k = ih[IH_KF + f], loc = p - offs[f], ic = loc / k, ia = loc - ic * k;
                const double* tc = C.tau_c + ((size_t)chain * T.nf + f) * C.kcap; const double* ta = C.tau_a + ((size_t)chain * T.nf + f) * C.kcap;
                const double* M = C.M + ((size_t)chain * (T.nf + 1) + ih[IH_MSLOT + f]) * ((size_t)C.kcap * C.ld);
                Mv = M[(size_t)ic * C.ld + ia];
                const double dtau = ta[ia] - tc[ic]; int idt;
                if (dtau > 0.0) { idt = (int)(dtau / beta * (double)R.ntau) + 1; xx = 2.0 / beta * dtau - 1.0; }
                else { idt = (int)((beta + dtau) / beta * (double)R.ntau) + 1; xx = 2.0 / beta * (beta + dtau) - 1.0; Mv = -Mv; }
                Mv *= cfgsign;
                if (idt > R.ntau) idt = R.ntau;
                atomic_add(&acc[A.o_gtau + f * R.ntau + idt - 1], Mv);
            }
            // lanes of one flavour are contiguous: segmented shuffle reduction, the first lane of each segment adds into the
            // warp's private accumulator (no atomics: shared-memory fp64 atomics are compare-and-swap loops)
            const int seg_first = act ? (offs[f] > base ? offs[f] - base : 0) : lane;
            const int seg_last = act ? (offs[f + 1] - 1 - base < W2B_WARP - 1 ? offs[f + 1] - 1 - base : W2B_WARP - 1) : lane;
            double* g = gleg_sh + f * R.nlegmax;
            double Plx = 0.0, Plm1 = 0.0, Pl = 0.0;
#pragma unroll 1
            for (int l = 0; l < R.nlegmax; ++l) {
                if (l == 0) Pl = 1.0; else if (l == 1) Pl = xx;
                else Pl = ldg(&T.leg_a[l]) * xx * Plx - ldg(&T.leg_b[l]) * Plm1;
                Plm1 = Plx; Plx = Pl;
                double v = Mv * Pl;
#if W2B_DEVICE
#pragma unroll
                for (int o = 1; o < W2B_WARP; o <<= 1) { const double t = __shfl_down_sync(W2B_FULL, v, o); if (lane + o <= seg_last) v += t; }
#endif
                if (act && lane == seg_first) g[l] += v;
            }
            wsync();
        }
        return;
    }
#pragma unroll 1
    for (int f = 0; f < T.nf; ++f) {
        const int k = ih[IH_KF + f];
        if (k == 0) continue;
        const double* tc = C.tau_c + ((size_t)chain * T.nf + f) * C.kcap; const double* ta = C.tau_a + ((size_t)chain * T.nf + f) * C.kcap;
        const double* M = C.M + ((size_t)chain * (T.nf + 1) + ih[IH_MSLOT + f]) * ((size_t)C.kcap * C.ld);
#pragma unroll 1
        for (int lbase = 0; lbase < R.nlegmax; lbase += LEG_CHUNK) {
            double part[LEG_CHUNK];                       // statically indexed below -> registers
#pragma unroll
            for (int q = 0; q < LEG_CHUNK; ++q) part[q] = 0.0;
#pragma unroll 1
            for (int p = lane; p < k * k; p += W2B_WARP) {
                int ic = p / k, ia = p - ic * k;
                double Mv = M[(size_t)ic * C.ld + ia];
                double dtau = ta[ia] - tc[ic], xx; int idt;
                if (dtau > 0.0) { idt = (int)(dtau / beta * (double)R.ntau) + 1; xx = 2.0 / beta * dtau - 1.0; }
                else { idt = (int)((beta + dtau) / beta * (double)R.ntau) + 1; xx = 2.0 / beta * (beta + dtau) - 1.0; Mv = -Mv; }
                Mv *= cfgsign;
                if (lbase == 0) { if (idt > R.ntau) idt = R.ntau; atomic_add(&acc[A.o_gtau + f * R.ntau + idt - 1], Mv); }
                // three-term recurrence (CTQMC.F90:2407-2420): advance to l = lbase, then accumulate LEG_CHUNK terms
                double Plx = 0.0, Plm1 = 0.0, Pl = 0.0;
#pragma unroll 1
                for (int l = 0; l < lbase; ++l) {
                    if (l == 0) Pl = 1.0; else if (l == 1) Pl = xx;
                    else { int ll = l - 1; Pl = (2.0 * ll + 1.0) / (double)(ll + 1) * xx * Plx - (double)ll / (double)(ll + 1) * Plm1; }
                    Plm1 = Plx; Plx = Pl;
                }
#pragma unroll
                for (int q = 0; q < LEG_CHUNK; ++q) {
                    const int l = lbase + q;
                    if (l == 0) Pl = 1.0; else if (l == 1) Pl = xx;
                    else { int ll = l - 1; Pl = (2.0 * ll + 1.0) / (double)(ll + 1) * xx * Plx - (double)ll / (double)(ll + 1) * Plm1; }
                    part[q] += Mv * Pl;
                    Plm1 = Plx; Plx = Pl;
                }
            }
#pragma unroll
            for (int q = 0; q < LEG_CHUNK; ++q) {
                double s = wsum(part[q]);
                if (lane == 0 && lbase + q < R.nlegmax) atomic_add(gleg_sh ? &gleg_sh[f * R.nlegmax + lbase + q] : &acc[A.o_gleg + f * R.nlegmax + lbase + q], s);
            }
        }
    }
}

// ------------------------------------------------------------------ Matsubara estimator (MeasGiw; CTQMC.F90:2400-2414, measure_giw_nfft 1015-1056)
// Giw(b, s, n) += sum over (annihilator a, creator c) pairs of the flavour:  -cfgsign M(c, a) exp(i w_n (tau_a - tau_c)),  w_n = (2n + 1) pi / beta,
// n = -Niw .. Niw - 1.  The reference evaluates these sums with NFFT (non-equispaced FFT, Nfft_z.F90:133-200, an approximation to ~1e-12);
// here they are summed DIRECTLY: one thread owns GIW_F consecutive frequencies, starts each pair with one sincospi (exact argument
// reduction) and steps through its frequencies with the pair's phase factor exp(2 pi i (tau_a - tau_c) / beta) -- 6 FP64 operations per
// (pair, frequency).  A CTA serves chains of ONE bin and keeps its partial sums in registers over all of them: one atomic per
// (thread, frequency) at the end instead of one per chain.
constexpr int GIW_F = 8;
W2B_HD void giw_accumulate_chain(const Tables& T, const Run& R, const Chains& C, int chain, int f, int m0, double* re, double* im) {
    const int* ih = C.ih + (size_t)chain * IH_STRIDE;
    const int k = ih[IH_KF + f];
    if (k == 0) return;
    const double ms = -C.dh[(size_t)chain * DH_STRIDE + DH_CFGSIGN];
    const double* tc = C.tau_c + ((size_t)chain * T.nf + f) * C.kcap; const double* ta = C.tau_a + ((size_t)chain * T.nf + f) * C.kcap;
    const double* M = C.M + ((size_t)chain * (T.nf + 1) + ih[IH_MSLOT + f]) * ((size_t)C.kcap * C.ld);
    const double ib = 1.0 / T.beta, odd = (double)(2 * (m0 - R.ngiw) + 1);
#pragma unroll 1
    for (int ic = 0; ic < k; ++ic) {
        const double t_c = tc[ic];
#pragma unroll 1
        for (int ia = 0; ia < k; ++ia) {
            const double x = (ta[ia] - t_c) * ib, mv = ms * M[(size_t)ic * C.ld + ia];
            double zs, zc, ss, sc;
            w_sincospi(odd * x, zs, zc);                     // exp(i w_{m0} (tau_a - tau_c))
            w_sincospi(2.0 * x, ss, sc);                     // step to the next frequency
#pragma unroll
            for (int q = 0; q < GIW_F; ++q) {
                re[q] += mv * zc; im[q] += mv * zs;
                const double nc = zc * sc - zs * ss; zs = zc * ss + zs * sc; zc = nc;
            }
        }
    }
}

// Rebuild all derived state of one chain from its operator list (set_mc_config_from_arrays, Trace.F90:6872-7000):
// superstate sequence, flavour tau arrays, inverse matrices + determinant, trace, signs.
W2B_HD void rebuild_chain(const Tables& T, const Run& R, const Chains& C, int chain, double* wscratch, int* iscratch, double* mscratch) {
    Ctx x;
    ctx_bind(T, R, C, chain, x, wscratch, iscratch, mscratch);
    const int lane = lane_id();
    x.wpos = 0; set_window_edges(T, R, x);
    if (lane == 0) for (int f = 0; f <= T.nf; ++f) x.ih[IH_MSLOT + f] = f;
    // flavour arrays
#pragma unroll 1
    for (int f = 0; f < T.nf; ++f) {
        int kc = 0, ka = 0;
        if (lane == 0) {
#pragma unroll 1
            for (int i = 0; i < x.n; ++i) {
                int cd = x.code[i];
                if ((cd >> 1) != f) continue;
                if ((cd & 1) == 0) { if (kc < C.kcap) x.tc[(size_t)f * C.kcap + kc] = x.tau[i]; ++kc; }
                else { if (ka < C.kcap) x.ta[(size_t)f * C.kcap + ka] = x.tau[i]; ++ka; }
            }
            if (kc != ka || kc > C.kcap) { x.err |= ERR_CAPACITY; kc = 0; }
            x.ih[IH_KF + f] = kc;
        }
        wsync();
    }
    x.err = wbcast(x.err, 0);
    int fin = walk_store(T, x.code, 0, x.n, x.outer, x.sst, 0);
    if (fin != x.outer) x.err |= ERR_GLOBAL_QN;
    double lg, sg;
    rebuild_inverses(T, C, x, x.tc, x.ta, x.ih + IH_KF, true, lg, sg);
    x.det_log = lg; x.det_sign = sg;
    if (sg == 0.0) x.err |= ERR_SINGULAR;
    if (!(x.err & ERR_GLOBAL_QN)) {
        if (T.dmax > 1) {
            if (x.n > 0) { full_trace(T, x, x.tau, x.code, x.n, x.outer, x.tr_log, x.tr_sign, x.only); commit_parts(T, x); }
            else { x.tr_log = R.statesampling ? -(T.evals[T.offset[x.outer] + x.ostate] - T.egs) * T.beta : T.empty_log[x.outer]; x.tr_sign = 1.0; empty_parts(T, x, x.outer); }
            slots_reset(C, x);
        } else trace_full_scalar(T, x.tau, x.code, x.sst, x.n, x.outer, x.tr_log, x.tr_sign);
    }
    x.npairs = -1;
    x.perm = perm_full(x.code, x.n);
    x.palt = 0; for (int f = 0; f < T.nf; ++f) x.palt |= palt_flavour(C, x, f) << f;
    update_cfgsign(T, x);
    ctx_store(C, chain, x);
}

}  // namespace w2b
