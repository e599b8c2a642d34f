// Host driver of the chain population: device memory layout, kernel launches, accumulator traffic.
//
// Memory layout in HBM (chain-major, every per-chain array contiguous so a warp's accesses coalesce):
//   op_tau  [chain][ops_cap] f64   op_code [chain][ops_cap] u8   op_sst [chain][ops_cap] i16
//   tau_c / tau_a [chain][nf][kcap] f64     (time-ordered creator / annihilator times per flavour)
//   M       [chain][nf+1][kcap][ld] f64     (inverse hybridisation matrices + one spare for out-of-place updates)
//   ih/dh   [chain][IH_STRIDE] i32 / [DH_STRIDE] f64   (scalar chain state), rng counters, move counters
//   accumulators [bin][stride] f64          (bin = global chain id mod nbins; reduced with one allreduce per block)
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "atomic.hpp"
#include "engine_core.hpp"

#if W2B_DEVICE
#include <cuda_runtime.h>
#define W2B_CUDA_CHECK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call); } while (0)
#endif

namespace w2b {

// ---------------------------------------------------------------- raw memory helpers (device or host-emulation)
template <class T_> T_* dev_alloc(size_t n) {
    if (n == 0) n = 1;
#if W2B_DEVICE
    void* p = nullptr; W2B_CUDA_CHECK(cudaMalloc(&p, n * sizeof(T_))); W2B_CUDA_CHECK(cudaMemset(p, 0, n * sizeof(T_))); return (T_*)p;
#else
    return (T_*)std::calloc(n, sizeof(T_));
#endif
}
template <class T_> void dev_free(T_*& p) {
    if (!p) return;
#if W2B_DEVICE
    cudaFree((void*)p);
#else
    std::free((void*)p);
#endif
    p = nullptr;
}
// bytes moved between host and device through these helpers (bench.py reports them per end-to-end step)
inline unsigned long long& h2d_bytes() { static unsigned long long v = 0; return v; }
inline unsigned long long& d2h_bytes() { static unsigned long long v = 0; return v; }
template <class T_> void h2d(T_* d, const T_* h, size_t n) {
    if (!n) return;
    h2d_bytes() += n * sizeof(T_);
#if W2B_DEVICE
    W2B_CUDA_CHECK(cudaMemcpy((void*)d, h, n * sizeof(T_), cudaMemcpyHostToDevice));
#else
    std::memcpy((void*)d, h, n * sizeof(T_));
#endif
}
template <class T_> void d2h(T_* h, const T_* d, size_t n) {
    if (!n) return;
    d2h_bytes() += n * sizeof(T_);
#if W2B_DEVICE
    W2B_CUDA_CHECK(cudaMemcpy(h, d, n * sizeof(T_), cudaMemcpyDeviceToHost));
#else
    std::memcpy(h, d, n * sizeof(T_));
#endif
}
template <class T_> void dev_zero(T_* d, size_t n) {
    if (!n) return;
#if W2B_DEVICE
    W2B_CUDA_CHECK(cudaMemset((void*)d, 0, n * sizeof(T_)));
#else
    std::memset((void*)d, 0, n * sizeof(T_));
#endif
}
template <class T_> T_* dev_upload(const std::vector<T_>& v) { T_* d = dev_alloc<T_>(v.size()); h2d(d, v.data(), v.size()); return d; }

// ---------------------------------------------------------------- kernels
constexpr int WARPS_PER_CTA = 4;
constexpr int MAX_CTA_THREADS = 1024;
// per-warp work space of the blocked trace in doubles: two d x m matrices, column log scales, per-state trace parts
#if W2B_DEVICE
__host__ __device__
#endif
inline int mat_doubles(const Chains& C, bool second) { return C.dmax > 1 ? C.xdoubles + 3 * mat_cols(C.dmax) + MAT_DMAX + (second ? C.mstride : 0) : 0; }

#if W2B_DEVICE
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_init(Tables T, Run R, Chains C) {
    int chain = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    if (chain < C.nchains) init_chain(T, R, C, chain);
}
#endif
#if W2B_DEVICE
#define W2B_HOSTDEV __host__ __device__
#else
#define W2B_HOSTDEV
#endif
// ---- persistent CTA-scheduler kernel (1x1 blocks)
W2B_HOSTDEV inline size_t sched_bytes(int ncb) { return (size_t)ncb * (sizeof(Prop) + sizeof(int) + 5 * sizeof(short) + 2) + 64; }
W2B_HOSTDEV inline void sched_carve(CtaShared& S, char* base, int ncb) {
    S.prop = (Prop*)base; base += (size_t)ncb * sizeof(Prop);
    S.left = (int*)base; base += (size_t)ncb * sizeof(int);
    S.lA = (short*)base; base += (size_t)ncb * sizeof(short);
    S.lR = (short*)base; base += (size_t)ncb * sizeof(short);
    S.lA2 = (short*)base; base += (size_t)ncb * sizeof(short);
    S.lR2 = (short*)base; base += (size_t)ncb * sizeof(short);
    S.lG = (short*)base; base += (size_t)ncb * sizeof(short);
    S.parked = (unsigned char*)base; base += ncb;
    S.kind = (unsigned char*)base;
}
#if W2B_DEVICE
// persistent sweep kernel: every CTA serves batches of ncb chains (batch b -> chains [b*ncb, (b+1)*ncb))
__global__ void __launch_bounds__(MAX_CTA_THREADS, 1) k_sweep(Tables T_, Run R_, Chains C_, Accum A_, long long nsteps, int wdoubles, int ncb, int nbatch,
                                                             int log_chain, double* logbuf) {
    extern __shared__ double smem[];
    __shared__ CtaShared S;
    // one chain context per warp and one copy of the parameter structs per CTA in shared memory (not one per thread in local memory)
    __shared__ Ctx ctxs[MAX_CTA_THREADS / 32];
    __shared__ Tables T; __shared__ Run R; __shared__ Chains C; __shared__ Accum A;
    if (threadIdx.x == 0) { T = T_; R = R_; C = C_; A = A_; }
    __syncthreads();
    const int W = blockDim.x >> 5, warp = threadIdx.x >> 5;
    double* ws = smem + (size_t)warp * wdoubles;
    int* is = (int*)(ws + 3 * C.kcap);
    if (threadIdx.x == 0) { CtaShared tmp; sched_carve(tmp, (char*)(smem + (size_t)W * wdoubles), ncb);
        S.prop = tmp.prop; S.left = tmp.left; S.lA = tmp.lA; S.lR = tmp.lR; S.lA2 = tmp.lA2; S.lR2 = tmp.lR2; S.lG = tmp.lG; S.parked = tmp.parked; S.kind = tmp.kind; }
    __syncthreads();
    for (int b = blockIdx.x; b < nbatch; b += gridDim.x) {
        const int lo = b * ncb, nloc = min(ncb, C.nchains - lo);
        cta_sweep(T, R, C, A, lo, nloc, nsteps, S, ws, is, ws + 3 * C.kcap, log_chain, logbuf, ctxs[warp]);
        __syncthreads();
    }
}
#endif
// ---- round kernels (blocked traces)
#ifndef W2B_PROPOSE_WARPS
#define W2B_PROPOSE_WARPS 2
#endif
#ifndef W2B_PROPOSE_OCC
#define W2B_PROPOSE_OCC 20
#endif
#ifndef W2B_DECIDE_OCC
#define W2B_DECIDE_OCC 8
#endif
constexpr int PROPOSE_WARPS = W2B_PROPOSE_WARPS;      // warps per CTA of the propose kernel (its queue pushes are aggregated per CTA)
#ifndef W2B_TRACE_OCC
#define W2B_TRACE_OCC 8
#endif
constexpr int TRACE_WARPS = 4;        // warps per CTA of the trace kernel (one work matrix per warp in shared memory)
constexpr int TICKET_BATCH = 1;       // tickets a warp takes per atomic on the round's ticket counter
constexpr int DECIDE_WARPS = 4;       // warps per CTA of the decide kernel (one work matrix per warp in shared memory)
#if W2B_DEVICE
__global__ void __launch_bounds__(256) k_sweep_begin(Chains C, long long nsteps) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C.nchains) { C.left[c] = (int)nsteps; C.parked[c] = 0; }
    if (c < 2 * SC_STRIDE) C.sched[c] = 0;
}
// ---- autonomous sweep kernel (1x1 blocks, large orders): warps take chains from one ticket counter and run them start to finish
constexpr int AUTO_WARPS = 8;
#ifndef W2B_AUTO_OCC
#define W2B_AUTO_OCC 4
#endif
__global__ void __launch_bounds__(AUTO_WARPS * 32, W2B_AUTO_OCC) k_sweep_auto(Tables T_, Run R_, Chains C_, Accum A_, long long nsteps, int wdoubles, int log_chain, double* logbuf) {
    extern __shared__ double smem[];
    __shared__ Ctx ctxs[AUTO_WARPS]; __shared__ Prop props[AUTO_WARPS];
    __shared__ Tables T; __shared__ Run R; __shared__ Chains C; __shared__ Accum A;
    if (threadIdx.x == 0) { T = T_; R = R_; C = C_; A = A_; }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    double* ws = smem + (size_t)warp * wdoubles;
    int* is = (int*)(ws + 3 * C.kcap);
    for (int chain = wgrab(&C.sched[SC_TICKET]); chain < C.nchains; chain = wgrab(&C.sched[SC_TICKET]))
        chain_sweep(T, R, C, A, chain, nsteps, props[warp], ws, is, ws + 3 * C.kcap, log_chain, logbuf, ctxs[warp]);
}
// propose phase of one round: one warp per chain; the CTA's results reach the round's queues with one atomic per queue
// (the parameter structs are copied to shared memory once per CTA: passing kernel parameters by reference to out-of-line
// functions would otherwise spill a private copy per thread to local memory)
__global__ void __launch_bounds__(PROPOSE_WARPS * 32, W2B_PROPOSE_OCC) k_propose(Tables T_, Run R_, Chains C_, long long nsteps, int parity, int log_chain, double* logbuf) {
    __shared__ ProposeOut outs[PROPOSE_WARPS];
    __shared__ Ctx ctxs[PROPOSE_WARPS];
    __shared__ Tables T; __shared__ Run R; __shared__ Chains C;
    if (threadIdx.x == 0) { T = T_; R = R_; C = C_; }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x < SC_STRIDE) C.sched[(parity ^ 1) * SC_STRIDE + threadIdx.x] = 0;     // counters of the next round
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) outs[warp] = ProposeOut{0, 0, -1, 0, -1, 0, 0, 0};
    __syncwarp();
    const int chain = blockIdx.x * PROPOSE_WARPS + warp;
    if (chain < C.nchains) phase_propose(T, R, C, chain, nsteps, parity, log_chain, logbuf, &outs[warp], ctxs[warp]);
    __syncthreads();
    if (warp == 0) propose_flush(C, parity, outs, PROPOSE_WARPS);
}
// decide phase of one round: persistent warps take tickets (rare moves first, then decide queues by outer block size)
__global__ void __launch_bounds__(DECIDE_WARPS * 32, W2B_DECIDE_OCC) k_decide(Tables T_, Run R_, Chains C_, Accum A_, long long nsteps, int wdoubles, int mdoubles, int parity,
                                                              int log_chain, double* logbuf) {
    extern __shared__ double smem[];
    __shared__ Tables T; __shared__ Run R; __shared__ Chains C; __shared__ Accum A;
    if (threadIdx.x == 0) { T = T_; R = R_; C = C_; A = A_; }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    double* ws = smem + (size_t)warp * wdoubles;
    int* is = (int*)(ws + 3 * C.kcap + mdoubles);
    int* sc = C.sched + parity * SC_STRIDE;
    __shared__ Ctx ctxs[DECIDE_WARPS];
    const int total = sc[SC_NRFIN] + sc[SC_NQ0] + sc[SC_NQ1] + sc[SC_NQ2];
#ifdef W2B_STATIC_TICKETS
    for (int t = blockIdx.x * DECIDE_WARPS + warp; t < total; t += gridDim.x * DECIDE_WARPS)
        phase_decide(T, R, C, A, t, parity, nsteps, log_chain, logbuf, ws, is, ws + 3 * C.kcap, ctxs[warp]);
#else
    for (int t0 = wgrab_n(&sc[SC_TICKET], TICKET_BATCH); t0 < total; t0 = wgrab_n(&sc[SC_TICKET], TICKET_BATCH))
      for (int t = t0; t < t0 + TICKET_BATCH && t < total; ++t)
        phase_decide(T, R, C, A, t, parity, nsteps, log_chain, logbuf, ws, is, ws + 3 * C.kcap, ctxs[warp]);
#endif
}
// global moves / tau shifts, on their own stream beside the round pipeline: scan the chain states, serve ST_RARE_NEW
__global__ void __launch_bounds__(DECIDE_WARPS * 32, 4) k_rare(Tables T_, Run R_, Chains C_, Accum A_, long long nsteps, int wdoubles, int mdoubles,
                                                              int log_chain, double* logbuf) {
    extern __shared__ double smem[];
    __shared__ Tables T; __shared__ Run R; __shared__ Chains C; __shared__ Accum A;
    if (threadIdx.x == 0) { T = T_; R = R_; C = C_; A = A_; }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    double* ws = smem + (size_t)warp * wdoubles;
    int* is = (int*)(ws + 3 * C.kcap + mdoubles);
    __shared__ Ctx ctxs[DECIDE_WARPS];
    // every warp scans 32 chain states per load (1 % of the chains have a rare move pending) and serves the flagged ones
    for (int base = (blockIdx.x * DECIDE_WARPS + warp) * 32; base < C.nchains; base += gridDim.x * DECIDE_WARPS * 32) {
        const int c0 = base + (threadIdx.x & 31);
        unsigned m = __ballot_sync(0xffffffffu, c0 < C.nchains && vload_u8(&C.parked[c0]) == ST_RARE_NEW);
        while (m) {
            const int c = base + __ffs(m) - 1; m &= m - 1;
            phase_rare(T, R, C, A, c, nsteps, log_chain, logbuf, ws, is, ws + 3 * C.kcap, ctxs[warp]);
        }
    }
}
// trace phase of one round (blocked traces): operator products of every surviving proposal on the FP64 tensor pipe
__global__ void __launch_bounds__(TRACE_WARPS * 32, W2B_TRACE_OCC) k_trace(Tables T, Chains C, int tdoubles, int parity) {
    extern __shared__ double smem[];
    double* ws = smem + (size_t)(threadIdx.x >> 5) * tdoubles;
    int* sc = C.sched + parity * SC_STRIDE;
    const int total = sc[SC_NT0] + sc[SC_NT1] + sc[SC_NT2];
#ifdef W2B_STATIC_TICKETS
    for (int t = blockIdx.x * TRACE_WARPS + (threadIdx.x >> 5); t < total; t += gridDim.x * TRACE_WARPS) phase_trace(T, C, t, parity, ws);
#else
    for (int t0 = wgrab_n(&sc[SC_TTICKET], TICKET_BATCH); t0 < total; t0 = wgrab_n(&sc[SC_TTICKET], TICKET_BATCH))
        for (int t = t0; t < t0 + TICKET_BATCH && t < total; ++t) phase_trace(T, C, t, parity, ws);
#endif
}
// set_taudiffmax -> init_window for every chain: window_position = 0, NPairs invalid (Trace.F90:1608-1648)
__global__ void __launch_bounds__(256) k_reset_window(Chains C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C.nchains) { C.ih[(size_t)c * IH_STRIDE + IH_WPOS] = 0; C.ih[(size_t)c * IH_STRIDE + IH_NPAIRS] = -1; }
}
// population health in three integers: OR of the chains' error flags, largest operator count, largest flavour order
__global__ void __launch_bounds__(256) k_flags(Chains C, int nf, int* out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    int e = 0, mo = 0, mk = 0;
    if (c < C.nchains) {
        const int* ih = C.ih + (size_t)c * IH_STRIDE;
        e = ih[IH_ERR]; mo = ih[IH_NOPS];
        for (int f = 0; f < nf; ++f) mk = max(mk, ih[IH_KF + f]);
    }
    for (int o = 16; o > 0; o >>= 1) { e |= __shfl_xor_sync(W2B_FULL, e, o); mo = max(mo, __shfl_xor_sync(W2B_FULL, mo, o)); mk = max(mk, __shfl_xor_sync(W2B_FULL, mk, o)); }
    if ((threadIdx.x & 31) == 0) { if (e) atomicOr(&out[0], e); atomicMax(&out[1], mo); atomicMax(&out[2], mk); }
}
// autocorrelation accumulator: one thread per (bin, component)
__global__ void __launch_bounds__(256) k_autocorr_add(const double* acc, long long stride, int o_gtau, int ncomp, int nbins, int nlevels, long long n,
                                                     const double* inv_group, double* prev, double* dat) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nbins * ncomp) return;
    const int bin = (int)(i / ncomp), comp = (int)(i % ncomp);
    const double now = acc[(size_t)bin * stride + o_gtau + comp];
    const double val = (now - prev[i]) * inv_group[bin];
    prev[i] = now;
    autocorr_add_one(val, n, nlevels, dat + (size_t)bin * nlevels * 3 * ncomp, ncomp, comp);
}
// measurement pre-pass (blocked traces): extend every chain's cache frontiers to tau = beta/2, one job per 8 outer states
__global__ void __launch_bounds__(TRACE_WARPS * 32, 6) k_meas_extend(Tables T, Chains C, int tdoubles) {
    extern __shared__ double smem[];
    double* ws = smem + (size_t)(threadIdx.x >> 5) * tdoubles;
    const int njobs = C.nchains * 4;
    for (int t = blockIdx.x * TRACE_WARPS + (threadIdx.x >> 5); t < njobs; t += gridDim.x * TRACE_WARPS)
        trace_job(T, C, (t >> 2) | ((t & 3) << Q_NTSHIFT) | Q_MEAS, ws);
}
// measurement: grid = nbins x nsub CTAs; CTA (bin, sub) serves every nsub-th chain of its bin and keeps the Legendre
// accumulator of the bin in shared memory (one flush per CTA instead of nf * NLegMax same-address atomics per chain)
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 2) k_measure(Tables T_, Run R_, Chains C_, Accum A_, int wdoubles, int mdoubles, int nsub) {
    extern __shared__ double smem[];
    __shared__ Tables T; __shared__ Run R; __shared__ Chains C; __shared__ Accum A;
    if (threadIdx.x == 0) { T = T_; R = R_; C = C_; A = A_; }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    double* ws = smem + (size_t)warp * wdoubles;
    int* is = (int*)(ws + 3 * C.kcap + mdoubles);
    double* gleg_sh = smem + (size_t)WARPS_PER_CTA * wdoubles;       // one accumulator per warp
    const int nleg = T.nf * R.nlegmax;
    for (int i = threadIdx.x; i < nleg * WARPS_PER_CTA; i += blockDim.x) gleg_sh[i] = 0.0;
    __syncthreads();
    const int bin = blockIdx.x % R.nbins, sub = blockIdx.x / R.nbins;
    const int c0 = (int)(((unsigned)bin + (unsigned)R.nbins - R.chain_base % (unsigned)R.nbins) % (unsigned)R.nbins);   // first chain of this bin
    for (int chain = c0 + R.nbins * (sub * WARPS_PER_CTA + warp); chain < C.nchains; chain += R.nbins * nsub * WARPS_PER_CTA)
        measure_chain(T, R, C, A, chain, ws, is, ws + 3 * C.kcap, gleg_sh + (size_t)warp * nleg);
    __syncthreads();
    double* acc = A.base + (size_t)bin * A.stride + A.o_gleg;
    for (int i = threadIdx.x; i < nleg; i += blockDim.x) {
        double v = 0.0;
        for (int w = 0; w < WARPS_PER_CTA; ++w) v += gleg_sh[(size_t)w * nleg + i];
        if (v != 0.0) atomicAdd(&acc[i], v);
    }
}
// Matsubara estimator (MeasGiw): grid = (nbins * nsub, nf, frequency tiles of GIW_THREADS * GIW_F); CTA (bin, sub) sums every nsub-th
// chain of its bin into register accumulators, one atomic per (thread, frequency) at the end
constexpr int GIW_THREADS = 128;
__global__ void __launch_bounds__(GIW_THREADS) k_meas_giw(Tables T, Run R, Chains C, Accum A, int nsub) {
    const int bin = blockIdx.x % R.nbins, sub = blockIdx.x / R.nbins, f = blockIdx.y;
    const int nfreq = 2 * R.ngiw, m0 = (blockIdx.z * GIW_THREADS + threadIdx.x) * GIW_F;
    if (m0 >= nfreq) return;
    double re[GIW_F], im[GIW_F];
#pragma unroll
    for (int q = 0; q < GIW_F; ++q) { re[q] = 0.0; im[q] = 0.0; }
    const int c0 = (int)(((unsigned)bin + (unsigned)R.nbins - R.chain_base % (unsigned)R.nbins) % (unsigned)R.nbins);   // first chain of this bin
    for (int chain = c0 + R.nbins * sub; chain < C.nchains; chain += R.nbins * nsub) giw_accumulate_chain(T, R, C, chain, f, m0, re, im);
    double* acc = A.base + (size_t)bin * A.stride + A.o_giw + 2 * ((size_t)f * nfreq + m0);
#pragma unroll
    for (int q = 0; q < GIW_F; ++q) if (m0 + q < nfreq) { atomicAdd(&acc[2 * q], re[q]); atomicAdd(&acc[2 * q + 1], im[q]); }
}
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_rebuild(Tables T, Run R, Chains C, int wdoubles, int mdoubles, int only_chain) {
    extern __shared__ double smem[];
    int warp = threadIdx.x >> 5;
    int chain = blockIdx.x * WARPS_PER_CTA + warp;
    if (chain >= C.nchains || (only_chain >= 0 && chain != only_chain)) return;
    double* ws = smem + (size_t)warp * wdoubles;
    int* is = (int*)(ws + 3 * C.kcap + mdoubles);
    rebuild_chain(T, R, C, chain, ws, is, ws + 3 * C.kcap);
}
#endif

struct EngineOptions {
    int nchains = 1024, nbins = 16, kcap = 0, ops_cap = 0;
    int cta_warps = 0, batch_chains = 0;      // sweep kernel geometry (0 = automatic)
    int scheduler = 0;                        // 1x1 blocks: 0 / 1 = persistent CTA-scheduler kernel, 2 = round kernels (blocked traces always use rounds)
    int lu_groups = 0;                        // (accepted and ignored: the shared-memory group LU of round 2 was replaced by invert_blocked)
    unsigned seed = 0, chain_base = 0;
};

class Engine {
public:
    AtomicTables at;
    Tables T{};
    Run R{};
    Chains C{};
    Accum A{};
    EngineOptions opt;
    long long accum_doubles = 0;       // per bin
    int wdoubles = 0, wdoubles_sweep = 0;   // per-warp scratch in doubles (measurement kernel: second work matrix; sweep kernel: one)
    long long launches = 0;            // kernels launched (bench accounting)
    int accpair_cap = 0;
    std::vector<void*> owned;
    std::vector<int> dens_off;         // per block: offset of its d x d slab in the density accumulator
    std::vector<double> ftau_stage;    // host staging of the diagonal hybridisation (w2b_upload_ftau)

    ~Engine() { release(); }

    void release() {
        for (void* p : owned) {
#if W2B_DEVICE
            cudaFree(p);
#else
            std::free(p);
#endif
        }
        owned.clear();
        d_flags = nullptr; ac_prev = ac_dat = ac_invg = nullptr; ac_levels = 0;
        release_host();
    }
    template <class T_> T_* keep(T_* p) { owned.push_back((void*)p); return p; }

    // ftau_diag: [nf][nftau] with flavour f = 2*band+spin
    std::vector<int> symmoves;         // [2 NSymMove][nf], set before setup()
    void setup(const AtomicTables& tables, const std::vector<double>& ftau_diag, int nftau, double beta, const Run& run,
               const EngineOptions& options) {
        release();
        at = tables; opt = options; R = run;
        if (at.nf > NF_MAX) throw std::runtime_error("too many flavours for this build (NF_MAX)");
        const int nf = at.nf, nsst = at.nsst;
        T.nb = at.nb; T.nf = nf; T.nsst = nsst; T.nops = at.nops; T.dmax = at.dmax; T.nftau = nftau; T.nstates = at.nstates;
        T.beta = beta; T.egs = at.egs;
        std::vector<short> conn(at.connect.begin(), at.connect.end());
        std::vector<double> e0(nsst), elem((size_t)nsst * at.nops, 0.0), ew(nsst), el(nsst);
        double wsum_ = 0;
        for (int k = 0; k < nsst; ++k) {
            e0[k] = at.evals[at.offset[k]];
            for (int c = 0; c < at.nops; ++c) if (at.op_off[(size_t)k * at.nops + c] >= 0) elem[(size_t)k * at.nops + c] = at.opdata[at.op_off[(size_t)k * at.nops + c]];
            // empty trace of outer superstate k: sum_j exp(-(E_j - Egs) beta)   (Trace.F90:647-655)
            double w = 0; for (int j = 0; j < at.dim[k]; ++j) w += std::exp(-(at.evals[at.offset[k] + j] - at.egs) * beta);
            ew[k] = w; el[k] = std::log(w); wsum_ += w;
        }
        for (double& w : ew) w /= wsum_;
        T.empty_norm = wsum_;
        std::vector<double> lelem(elem.size(), 0.0);
        for (size_t i = 0; i < elem.size(); ++i) if (elem[i] != 0.0) lelem[i] = std::log(std::fabs(elem[i]));
        T.connect = keep(dev_upload(conn)); T.e0 = keep(dev_upload(e0)); T.elem = keep(dev_upload(elem)); T.lelem = keep(dev_upload(lelem));
        T.ftau = keep(dev_upload(ftau_diag));
        T.first_state = keep(dev_upload(at.first_state)); T.state2sst = keep(dev_upload(at.state2sst));
        T.dim = keep(dev_upload(at.dim)); T.offset = keep(dev_upload(at.offset));
        std::vector<int> doff(nsst); int dsum = 0;
        for (int k = 0; k < nsst; ++k) { doff[k] = dsum; dsum += at.dim[k] * at.dim[k]; }
        dens_off = doff;
        T.dens_off = keep(dev_upload(doff));
        T.empty_w = keep(dev_upload(ew)); T.empty_log = keep(dev_upload(el));
        {
            std::vector<OpDesc> desc((size_t)nsst * at.nops);
            std::vector<double> frag;
            for (int k = 0; k < nsst; ++k) for (int c = 0; c < at.nops; ++c) {
                OpDesc& D = desc[(size_t)k * at.nops + c];
                const int out = at.connect[(size_t)k * at.nops + c];
                D.sst_out = out; D.d_in = at.dim[k]; D.d_out = out >= 0 ? at.dim[out] : 0; D.ev_off = at.offset[k];
                D.op_off = out >= 0 ? (int)at.op_off[(size_t)k * at.nops + c] : 0; D.e0 = at.evals[at.offset[k]];
                D.frag_off = (int)frag.size();
                if (out < 0 || at.dmax == 1) continue;
                // DMMA A-fragment order: [k-step][row-tile pair][lane][2]; lane l holds rows 8*mt + l/4, column 4*ks + l%4
                const int din = D.d_in, dout = D.d_out, ks_n = (din + 3) / 4, np = ((dout + 7) / 8 + 1) / 2;
                const double* O = at.opdata.data() + D.op_off;
                for (int ks = 0; ks < ks_n; ++ks) for (int p = 0; p < np; ++p) for (int l = 0; l < 32; ++l) for (int h = 0; h < 2; ++h) {
                    const int r = (2 * p + h) * 8 + l / 4, cc = ks * 4 + l % 4;
                    frag.push_back((r < dout && cc < din) ? O[r + (size_t)cc * dout] : 0.0);
                }
            }
            if (frag.empty()) frag.push_back(0.0);
            T.desc = keep(dev_upload(desc)); T.opfrag = keep(dev_upload(frag));
        }
        {
            std::vector<double> la(std::max(R.nlegmax, 2), 0.0), lb(std::max(R.nlegmax, 2), 0.0);
            for (int l = 2; l < R.nlegmax; ++l) { const int ll = l - 1; la[l] = (2.0 * ll + 1.0) / (double)(ll + 1); lb[l] = (double)ll / (double)(ll + 1); }
            T.leg_a = keep(dev_upload(la)); T.leg_b = keep(dev_upload(lb));
            std::vector<double> lg(la.size(), 0.0), lc(la.size(), 1.0);
            for (int l = 1; l < (int)lc.size(); ++l) lc[l] = lc[l - 1] * (2.0 * l - 1.0) / (double)l;       // c_l = (2l)! / (2^l l!^2)
            for (int l = 2; l < (int)lg.size(); ++l) lg[l] = ((double)(l - 1) * (l - 1)) / ((2.0 * l - 1.0) * (2.0 * l - 3.0));
            T.leg_g = keep(dev_upload(lg)); T.leg_c = keep(dev_upload(lc));
        }
        { std::vector<int> sm = symmoves; if (sm.empty()) sm.push_back(0); T.symmoves = keep(dev_upload(sm)); }
        T.op_off = keep(dev_upload(at.op_off)); T.opdata = keep(dev_upload(at.opdata)); T.evals = keep(dev_upload(at.evals));

        // capacities
        int kcap = opt.kcap;
        if (kcap <= 0) { double kd = 0.45 * beta; kcap = (int)(kd + 6.0 * std::sqrt(kd) + 16.0); }
        kcap = ((kcap + 15) / 16) * 16;
        int ops_cap = opt.ops_cap;
        if (ops_cap <= 0) {                      // 2 * (<n_pairs> + 8 sigma + 32); blocked traces keep one cache slot per operator
            double np = nf * 0.45 * beta; ops_cap = (int)(2.0 * (np + 8.0 * std::sqrt(np) + 32.0));
            if (at.dmax > 1 && ops_cap > 256) ops_cap = 256;
        }
        if (ops_cap > 2 * nf * kcap) ops_cap = 2 * nf * kcap;
        ops_cap = ((ops_cap + 31) / 32) * 32;
        if (ops_cap > 4096) throw std::runtime_error("ops_cap > 4096 not supported (16-bit cache slot ids)");
        const int nc = opt.nchains;
        C.nchains = nc; C.ops_cap = ops_cap; C.kcap = kcap; C.ld = kcap;
        C.op_tau = keep(dev_alloc<double>((size_t)nc * ops_cap)); C.op_code = keep(dev_alloc<unsigned char>((size_t)nc * ops_cap));
        C.op_sst = keep(dev_alloc<short>((size_t)nc * ops_cap));
        C.sc_tau = keep(dev_alloc<double>((size_t)nc * ops_cap)); C.sc_code = keep(dev_alloc<unsigned char>((size_t)nc * ops_cap));
        C.sc_sst = keep(dev_alloc<short>((size_t)nc * ops_cap));
        C.tau_c = keep(dev_alloc<double>((size_t)nc * nf * kcap)); C.tau_a = keep(dev_alloc<double>((size_t)nc * nf * kcap));
        C.sc_tc = keep(dev_alloc<double>((size_t)nc * nf * kcap)); C.sc_ta = keep(dev_alloc<double>((size_t)nc * nf * kcap));
        C.M = keep(dev_alloc<double>((size_t)nc * (nf + 1) * kcap * kcap));
        C.ih = keep(dev_alloc<int>((size_t)nc * IH_STRIDE)); C.dh = keep(dev_alloc<double>((size_t)nc * DH_STRIDE));
        C.rng_ctr = keep(dev_alloc<unsigned long long>(nc)); C.counters = keep(dev_alloc<unsigned long long>((size_t)nc * C_NCNT));
        C.dmax = at.dmax; C.mstride = at.dmax > 1 ? at.dmax * at.dmax : 0;
        if (at.dmax > MAT_DMAX) throw std::runtime_error("quantum-number blocks larger than 32 are not supported by the tensor trace path (add quantum numbers)");
        C.ss = R.statesampling;
        C.ldx = mat_ld(at.dmax); C.xdoubles = at.dmax > 1 ? C.ldx * mat_cols(at.dmax) : 0;
        if (at.dmax > 1) {
            C.op_slot = keep(dev_alloc<unsigned short>((size_t)nc * ops_cap)); C.free_slot = keep(dev_alloc<unsigned short>((size_t)nc * ops_cap));
            C.ket = keep(dev_alloc<double>((size_t)nc * ops_cap * C.mstride)); C.bra = keep(dev_alloc<double>((size_t)nc * ops_cap * C.mstride));
            C.ketlg = keep(dev_alloc<double>((size_t)nc * ops_cap * C.dmax)); C.bralg = keep(dev_alloc<double>((size_t)nc * ops_cap * C.dmax));
            C.pt = keep(dev_alloc<double>((size_t)nc * 2 * C.dmax));
        }
        C.left = keep(dev_alloc<int>(nc)); C.parked = keep(dev_alloc<unsigned char>(nc)); C.kind = keep(dev_alloc<unsigned char>(nc));
        C.prop = keep(dev_alloc<Prop>(nc)); C.rjob = keep(dev_alloc<RareJob>(nc));
        C.q_rfin = keep(dev_alloc<int>((size_t)2 * nc)); C.q_dec0 = keep(dev_alloc<int>(nc)); C.q_dec1 = keep(dev_alloc<int>(nc)); C.q_dec2 = keep(dev_alloc<int>(nc));
        C.sched = keep(dev_alloc<int>(2 * SC_STRIDE));
        C.ptc = nullptr; C.q_tr0 = C.q_tr1 = C.q_tr2 = nullptr;
        if (at.dmax > 1) {
            C.ptc = keep(dev_alloc<double>((size_t)nc * 2 * at.dmax));
            C.q_tr0 = keep(dev_alloc<int>((size_t)nc * 8)); C.q_tr1 = keep(dev_alloc<int>((size_t)nc * 8)); C.q_tr2 = keep(dev_alloc<int>((size_t)nc * 8));
        }
        int wints = kcap + 32 + 2 * nsst;      // + candidate outer superstates of a global move: new [nsst], old [nsst]
        wdoubles = 3 * kcap + mat_doubles(C, true) + (wints + 1) / 2;
        wdoubles_sweep = 3 * kcap + (wints + 1) / 2;   // sweep kernels hold no work matrix: blocked traces run in the trace kernel

        // accumulators
        int o = 0;
        A.o_gtau = o; o += nf * R.ntau; A.o_gleg = o; o += nf * R.nlegmax; A.o_histo = o; o += nf * (R.maxhisto + 1);
        A.ndens = dsum; A.o_dens = o; o += A.ndens;
        A.o_outer = o; o += nsst; A.o_signsst = o; o += nsst; A.o_contribsst = o; o += nsst;
        A.o_outerst = o; o += at.nstates; A.o_signst = o; o += at.nstates; A.o_contribst = o; o += at.nstates;
        A.o_sign = o; o += 1; A.o_cnt = o; o += 1;
        A.o_giw = o; o += 2 * nf * 2 * R.ngiw;
        accum_doubles = o; A.stride = o;
        A.base = keep(dev_alloc<double>((size_t)R.nbins * o));
        accpair_cap = std::max(R.ntau, 1024);
        A.accpair = keep(dev_alloc<unsigned long long>((size_t)nf * accpair_cap));
        R.naccpair = R.ntau; R.accpair_max = beta; R.freeze_window = 0;
        pairlist_alloc = 4096;
        A.pairlist = keep(dev_alloc<double>(pairlist_alloc)); A.pairlist_cnt = keep(dev_alloc<int>(1)); A.pairlist_cap = 0;
        init_chains();
    }

    size_t smem_bytes() const { return (size_t)WARPS_PER_CTA * wdoubles * sizeof(double); }
    int grid() const { return (C.nchains + WARPS_PER_CTA - 1) / WARPS_PER_CTA; }

    void sync() {
#if W2B_DEVICE
        W2B_CUDA_CHECK(cudaDeviceSynchronize());
#endif
    }

    void init_chains() {
#if W2B_DEVICE
        k_init<<<grid(), WARPS_PER_CTA * 32>>>(T, R, C); ++launches;
        W2B_CUDA_CHECK(cudaGetLastError());
#else
        for (int c = 0; c < C.nchains; ++c) init_chain(T, R, C, c);
#endif
    }

    void set_taudiffmax(double tdm) {          // set_taudiffmax -> init_window (CTQMC.F90:8180, Trace.F90:1608-1648)
        if (tdm > 0.0) { R.taudiff_max = tdm; R.num_windows = std::max((int)std::floor(T.beta / tdm) - 1, 1); }
        else { R.taudiff_max = T.beta; R.num_windows = 1; }
        // window_position = 0, NPairs invalid for every chain
#if W2B_DEVICE
        k_reset_window<<<(C.nchains + 255) / 256, 256>>>(C); ++launches;
        W2B_CUDA_CHECK(cudaGetLastError());
        return;
#endif
        std::vector<int> ih((size_t)C.nchains * IH_STRIDE);
        d2h(ih.data(), C.ih, ih.size());
        for (int c = 0; c < C.nchains; ++c) { ih[(size_t)c * IH_STRIDE + IH_WPOS] = 0; ih[(size_t)c * IH_STRIDE + IH_NPAIRS] = -1; }
        h2d(C.ih, ih.data(), ih.size());
    }

    void reset_counters() {                    // init_counters CTQMC.F90:847-899
        dev_zero(C.counters, (size_t)C.nchains * C_NCNT);
        dev_zero(A.accpair, (size_t)T.nf * accpair_cap);
    }
    void set_accpair_binning(int n, double range) {
        if (n > accpair_cap) throw std::runtime_error("accepted-pair histogram: too many bins");
        R.naccpair = n; R.accpair_max = range; dev_zero(A.accpair, (size_t)T.nf * accpair_cap);
    }
    // list of the tau distances of the first `n` accepted pair moves (ctqmc_calibrate phase 1); n = 0 switches it off
    int pairlist_alloc = 0;
    void pairlist_begin(int n) {
        if (n > pairlist_alloc) { A.pairlist = keep(dev_alloc<double>(n)); pairlist_alloc = n; }
        dev_zero(A.pairlist_cnt, 1); A.pairlist_cap = n;
    }
    int pairlist_count() { sync(); int c = 0; d2h(&c, A.pairlist_cnt, 1); return c; }
    void pairlist_get(std::vector<double>& out, int n) { sync(); out.resize(n); d2h(out.data(), A.pairlist, n); }
    void get_accpair(std::vector<unsigned long long>& out) { sync(); out.resize((size_t)T.nf * R.naccpair); d2h(out.data(), A.accpair, out.size()); }
    void zero_accum() { dev_zero(A.base, (size_t)R.nbins * accum_doubles); autocorr_reset(); }

    // ---- autocorrelation accumulator of G(tau) (AccumGtau = 2, Accumulators.F90:214-364), one series per chain-group bin
    int ac_levels = 0; long long ac_n = 0;
    double* ac_prev = nullptr; double* ac_dat = nullptr; double* ac_invg = nullptr;     // [bin][comp], [bin][level][3][comp], [bin]
    std::vector<double> ac_group;                                                       // chains per bin
    void autocorr_setup(int nlevels) {
        ac_levels = std::max(nlevels, 1);
        const size_t ncomp = (size_t)T.nf * R.ntau;
        ac_prev = keep(dev_alloc<double>((size_t)R.nbins * ncomp));
        ac_dat = keep(dev_alloc<double>((size_t)R.nbins * ac_levels * 3 * ncomp));
        ac_group.assign(R.nbins, 0.0);
        for (int c = 0; c < C.nchains; ++c) ac_group[(R.chain_base + (unsigned)c) % (unsigned)R.nbins] += 1.0;
        std::vector<double> inv(R.nbins);
        for (int b = 0; b < R.nbins; ++b) inv[b] = ac_group[b] > 0 ? 1.0 / ac_group[b] : 0.0;
        ac_invg = keep(dev_upload(inv));
        ac_n = 0;
    }
    void autocorr_reset() {
        if (!ac_levels) return;
        const size_t ncomp = (size_t)T.nf * R.ntau;
        dev_zero(ac_prev, (size_t)R.nbins * ncomp); dev_zero(ac_dat, (size_t)R.nbins * ac_levels * 3 * ncomp);
        ac_n = 0;
    }
    void autocorr_add() {                       // call after every measurement
        if (!ac_levels) return;
        const int ncomp = T.nf * R.ntau;
#if W2B_DEVICE
        const long long tot = (long long)R.nbins * ncomp;
        k_autocorr_add<<<(unsigned)((tot + 255) / 256), 256>>>(A.base, A.stride, A.o_gtau, ncomp, R.nbins, ac_levels, ac_n, ac_invg, ac_prev, ac_dat); ++launches;
        W2B_CUDA_CHECK(cudaGetLastError());
#else
        for (int b = 0; b < R.nbins; ++b) for (int comp = 0; comp < ncomp; ++comp) {
            const size_t i = (size_t)b * ncomp + comp;
            const double nowv = A.base[(size_t)b * A.stride + A.o_gtau + comp];
            const double val = (nowv - ac_prev[i]) * ac_invg[b];
            ac_prev[i] = nowv;
            autocorr_add_one(val, ac_n, ac_levels, ac_dat + (size_t)b * ac_levels * 3 * ncomp, ncomp, comp);
        }
#endif
        ++ac_n;
    }
    // autocorracc_blocks (Accumulators.F90:343-361): variance of the level-l bin means, here of a single chain's series
    // (group-mean variance x chains per bin, chains being independent), averaged over the bins; [level][comp]
    void autocorr_blocks(std::vector<double>& out) {
        const size_t ncomp = (size_t)T.nf * R.ntau;
        out.assign((size_t)ac_levels * ncomp, 0.0);
        if (!ac_levels) return;
        sync();
        std::vector<double> dat((size_t)R.nbins * ac_levels * 3 * ncomp);
        d2h(dat.data(), ac_dat, dat.size());
        int used = 0; for (int b = 0; b < R.nbins; ++b) used += ac_group[b] > 0;
        long long n = ac_n;
        for (int l = 0; l < ac_levels; ++l) {
            for (int b = 0; b < R.nbins; ++b) {
                if (!(ac_group[b] > 0)) continue;
                const double* sq = dat.data() + ((size_t)b * ac_levels + l) * 3 * ncomp + 2 * ncomp;
                for (size_t c = 0; c < ncomp; ++c) out[(size_t)l * ncomp + c] += (n > 1 ? sq[c] / (double)(n - 1) : sq[c] / 0.0) * ac_group[b] / used;
            }
            n /= 2;
        }
    }

    // ---- 1x1 blocks: persistent CTA-scheduler kernel; geometry: W warps per CTA, batches of ncb chains, persistent grid
    int ct_warps = 1, ct_ncb = 1, ct_nbatch = 1, ct_grid = 1; size_t ct_smem = 0; bool ct_ready = false;
    void plan_cta() {
        const int nc = C.nchains;
#if W2B_DEVICE
        int dev = 0, nsm = 1, smem_max = 48 * 1024;
        W2B_CUDA_CHECK(cudaGetDevice(&dev));
        W2B_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
        W2B_CUDA_CHECK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        cudaFuncAttributes fa; W2B_CUDA_CHECK(cudaFuncGetAttributes(&fa, k_sweep));
        smem_max -= (int)fa.sharedSizeBytes;                 // static shared memory of the kernel (contexts, parameter copies)
        int W = opt.cta_warps > 0 ? opt.cta_warps : 32;
        const size_t per_warp = (size_t)wdoubles_sweep * sizeof(double);
        while (W > 1 && W * per_warp + sched_bytes(8 * W) > (size_t)smem_max - 1024) --W;
        if (W > MAX_CTA_THREADS / 32) W = MAX_CTA_THREADS / 32;
        int target = opt.batch_chains > 0 ? opt.batch_chains : 8 * W;
        size_t sm = W * per_warp + sched_bytes(target);
        W2B_CUDA_CHECK(cudaFuncSetAttribute(k_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::min((size_t)smem_max, sm + 4096)));
        int occ = 1;
        W2B_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep, W * 32, sm));
        if (occ < 1) occ = 1;
        const int slots = nsm * occ;
        int rounds = (nc + slots * target - 1) / (slots * target);
        int ncb = (nc + slots * rounds - 1) / (slots * rounds);
        if (ncb < 1) ncb = 1;
        if (ncb > target) ncb = target;
        ct_warps = W; ct_ncb = ncb; ct_nbatch = (nc + ncb - 1) / ncb; ct_grid = std::min(ct_nbatch, slots);
        ct_smem = W * per_warp + sched_bytes(ncb);
#else
        ct_warps = 1; ct_ncb = opt.batch_chains > 0 ? opt.batch_chains : 4; ct_nbatch = (nc + ct_ncb - 1) / ct_ncb; ct_grid = 1;
#endif
        ct_ready = true;
    }

    // autonomous form (scheduler 3; automatic for 1x1 blocks when the capacity says the orders are large): one warp per chain, no phases
    int au_grid = 1; size_t au_smem = 0; bool au_ready = false;
    void sweep_auto(long long nsteps, int log_chain, double* log_host) {
        double* dlog = nullptr;
        if (log_chain >= 0) dlog = dev_alloc<double>((size_t)nsteps * LOGW);
#if W2B_DEVICE
        if (!au_ready) {
            int dev = 0, nsm = 1;
            W2B_CUDA_CHECK(cudaGetDevice(&dev));
            W2B_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
            au_smem = (size_t)AUTO_WARPS * wdoubles_sweep * sizeof(double);
            W2B_CUDA_CHECK(cudaFuncSetAttribute(k_sweep_auto, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)au_smem));
            int occ = 1;
            W2B_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep_auto, AUTO_WARPS * 32, au_smem));
            occ = std::min(occ, W2B_AUTO_OCC);
            au_grid = std::max(1, std::min(nsm * std::max(1, occ), (C.nchains + AUTO_WARPS - 1) / AUTO_WARPS));
            au_ready = true;
        }
        k_sweep_begin<<<(std::max(C.nchains, 2 * SC_STRIDE) + 255) / 256, 256>>>(C, nsteps); ++launches;        // zeroes the ticket counter
        k_sweep_auto<<<au_grid, AUTO_WARPS * 32, au_smem>>>(T, R, C, A, nsteps, wdoubles_sweep, log_chain, dlog); ++launches;
        W2B_CUDA_CHECK(cudaGetLastError());
#else
        std::vector<double> ws(wdoubles + 8);
        for (int c = 0; c < C.nchains; ++c) { Ctx x; Prop P;
            chain_sweep(T, R, C, A, c, nsteps, P, ws.data(), (int*)(ws.data() + 3 * C.kcap + mat_doubles(C, true)), ws.data() + 3 * C.kcap, log_chain, dlog, x); }
#endif
        if (log_chain >= 0) { sync(); d2h(log_host, dlog, (size_t)nsteps * LOGW); dev_free(dlog); }
    }

    // advance every chain by nsteps moves; optionally record the replay log of one chain
    void sweep_cta(long long nsteps, int log_chain, double* log_host) {
        if (!ct_ready) plan_cta();
        double* dlog = nullptr;
        if (log_chain >= 0) dlog = dev_alloc<double>((size_t)nsteps * LOGW);
#if W2B_DEVICE
        k_sweep<<<ct_grid, ct_warps * 32, ct_smem>>>(T, R, C, A, nsteps, wdoubles_sweep, ct_ncb, ct_nbatch, log_chain, dlog); ++launches;
        W2B_CUDA_CHECK(cudaGetLastError());
#else
        std::vector<double> ws(wdoubles + 8);
        std::vector<char> sb(sched_bytes(ct_ncb));
        CtaShared S{}; sched_carve(S, sb.data(), ct_ncb);
        for (int b = 0; b < ct_nbatch; ++b) {
            const int lo = b * ct_ncb, nloc = std::min(ct_ncb, C.nchains - lo);
            { Ctx x;
              cta_sweep(T, R, C, A, lo, nloc, nsteps, S, ws.data(), (int*)(ws.data() + 3 * C.kcap + mat_doubles(C, true)), ws.data() + 3 * C.kcap, log_chain, dlog, x); }
        }
#endif
        if (log_chain >= 0) { sync(); d2h(log_host, dlog, (size_t)nsteps * LOGW); dev_free(dlog); }
    }


    // ---- blocked traces: round kernels
    // decide-kernel geometry: persistent grid of DECIDE_WARPS-warp CTAs, as many as fit per SM
    int sw_warps = DECIDE_WARPS, sw_ncb = 0, sw_nbatch = 0, sw_grid = 1; size_t sw_smem = 0; bool sw_ready = false;
    int* h_sched = nullptr;            // pinned host copy of one counter set
#if W2B_DEVICE
    cudaStream_t rare_stream = nullptr; cudaEvent_t ev_prop[2] = {nullptr, nullptr}, ev_rare[2] = {nullptr, nullptr};
#endif
    size_t smem_cfg_measure = 0, smem_cfg_rebuild = 0;    // dynamic shared memory the measure / rebuild kernels are opted in for (this engine's device)
    int tr_doubles = 0, tr_grid = 1, rp_grid = 1; size_t tr_smem = 0;   // trace kernel: per-warp work space, grid, shared memory
    long long rounds = 0;              // rounds launched so far (bench accounting)
    bool time_kernels = false;         // bench: CUDA events around every round kernel (on the launching stream)
    double ms_propose = 0, ms_trace = 0, ms_decide = 0;
#if W2B_DEVICE
    std::vector<cudaEvent_t> tev;      // 4 events per round of a chunk
#endif
    unsigned long long trace_flops() { std::vector<unsigned long long> c; get_counters(c); return c[C_TRFLOPS]; }
    // streams, events and pinned staging of the round scheduler
    void release_host() {
#if W2B_DEVICE
        if (rare_stream) { cudaStreamDestroy(rare_stream); rare_stream = nullptr; }
        for (int i = 0; i < 2; ++i) { if (ev_prop[i]) cudaEventDestroy(ev_prop[i]); if (ev_rare[i]) cudaEventDestroy(ev_rare[i]); ev_prop[i] = ev_rare[i] = nullptr; }
        for (cudaEvent_t e : tev) cudaEventDestroy(e);
        tev.clear();
        if (h_sched) { cudaFreeHost(h_sched); h_sched = nullptr; }
#else
        if (h_sched) { std::free(h_sched); h_sched = nullptr; }
#endif
        sw_ready = false; ct_ready = false; au_ready = false; smem_cfg_measure = smem_cfg_rebuild = 0;      // the opt-in is per device: redo after a device change
    }
    void plan_sweep() {
#if W2B_DEVICE
        int dev = 0, nsm = 1;
        W2B_CUDA_CHECK(cudaGetDevice(&dev));
        W2B_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
        sw_smem = (size_t)DECIDE_WARPS * wdoubles_sweep * sizeof(double);
        W2B_CUDA_CHECK(cudaFuncSetAttribute(k_decide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sw_smem));
        int occ = 1;
        W2B_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_decide, DECIDE_WARPS * 32, sw_smem));
        if (occ < 1) occ = 1;
        if (opt.cta_warps > 0) occ = std::max(1, std::min(occ, opt.cta_warps / DECIDE_WARPS));     // option: cap resident decide warps per SM
        sw_grid = std::min(nsm * occ, (C.nchains + DECIDE_WARPS - 1) / DECIDE_WARPS);
        sw_warps = occ * DECIDE_WARPS;
        if (!h_sched) W2B_CUDA_CHECK(cudaMallocHost((void**)&h_sched, SC_STRIDE * sizeof(int)));
        if (!rare_stream) {
            W2B_CUDA_CHECK(cudaStreamCreateWithFlags(&rare_stream, cudaStreamNonBlocking));
            for (int i = 0; i < 2; ++i) { W2B_CUDA_CHECK(cudaEventCreateWithFlags(&ev_prop[i], cudaEventDisableTiming)); W2B_CUDA_CHECK(cudaEventCreateWithFlags(&ev_rare[i], cudaEventDisableTiming)); }
        }
        W2B_CUDA_CHECK(cudaFuncSetAttribute(k_rare, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sw_smem));
        rp_grid = std::max(1, std::min(nsm, C.nchains / (8 * DECIDE_WARPS) + 1));
        if (C.dmax > 1) {
            tr_doubles = trace_ws_doubles(C.ldx);
            tr_smem = (size_t)TRACE_WARPS * tr_doubles * sizeof(double);
            W2B_CUDA_CHECK(cudaFuncSetAttribute(k_trace, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tr_smem));
            W2B_CUDA_CHECK(cudaFuncSetAttribute(k_meas_extend, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tr_smem));
            int occt = 1;
            W2B_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occt, k_trace, TRACE_WARPS * 32, tr_smem));
            tr_grid = std::min(nsm * std::max(1, occt), (C.nchains + TRACE_WARPS - 1) / TRACE_WARPS);

        }
#else
        if (!h_sched) h_sched = (int*)std::calloc(SC_STRIDE, sizeof(int));
#endif
        sw_ready = true;
    }

    // advance every chain by nsteps moves; optionally record the replay log of one chain
    void sweep(long long nsteps, int log_chain = -1, double* log_host = nullptr) {
        if (nsteps <= 0) return;
        if (C.dmax == 1 && opt.scheduler == 3) { sweep_auto(nsteps, log_chain, log_host); return; }
        if (C.dmax == 1 && opt.scheduler != 2) { sweep_cta(nsteps, log_chain, log_host); return; }
        if (!sw_ready) plan_sweep();
        double* dlog = nullptr;
        if (log_chain >= 0) dlog = dev_alloc<double>((size_t)nsteps * LOGW);
        const int nc = C.nchains, prun = std::max(1, R.propose_run);
        long long chunk = std::max<long long>(1, nsteps / prun), round = 0;
#if W2B_DEVICE
        k_sweep_begin<<<(std::max(nc, 2 * SC_STRIDE) + 255) / 256, 256>>>(C, nsteps); ++launches;
        const int gp = (nc + PROPOSE_WARPS - 1) / PROPOSE_WARPS;
        const int md = 0;
#else
        for (int c = 0; c < nc; ++c) { C.left[c] = (int)nsteps; C.parked[c] = 0; }
        for (int i = 0; i < 2 * SC_STRIDE; ++i) C.sched[i] = 0;
        std::vector<double> ws(wdoubles + 8);
#endif
        for (;;) {
#if W2B_DEVICE
            if (time_kernels) {
                if (chunk > 64) chunk = 64;
                while (tev.size() < (size_t)chunk * 4) { cudaEvent_t e; W2B_CUDA_CHECK(cudaEventCreate(&e)); tev.push_back(e); }
            }
#endif
            for (long long i = 0; i < chunk; ++i, ++round) {
                const int parity = (int)(round & 1);
#if W2B_DEVICE
                if (time_kernels) W2B_CUDA_CHECK(cudaEventRecord(tev[4 * i], 0));
                k_propose<<<gp, PROPOSE_WARPS * 32>>>(T, R, C, nsteps, parity, log_chain, dlog);
                if (time_kernels) W2B_CUDA_CHECK(cudaEventRecord(tev[4 * i + 1], 0));
                // global moves / tau shifts drawn in this round: served on the rare stream, never waited for by the rounds
                W2B_CUDA_CHECK(cudaEventRecord(ev_prop[parity], 0));
                W2B_CUDA_CHECK(cudaStreamWaitEvent(rare_stream, ev_prop[parity], 0));
                k_rare<<<rp_grid, DECIDE_WARPS * 32, sw_smem, rare_stream>>>(T, R, C, A, nsteps, wdoubles_sweep, md, log_chain, dlog);
                ++launches;
                if (C.dmax > 1) { k_trace<<<tr_grid, TRACE_WARPS * 32, tr_smem>>>(T, C, tr_doubles, parity); ++launches; }
                if (time_kernels) W2B_CUDA_CHECK(cudaEventRecord(tev[4 * i + 2], 0));
                k_decide<<<sw_grid, DECIDE_WARPS * 32, sw_smem>>>(T, R, C, A, nsteps, wdoubles_sweep, md, parity, log_chain, dlog);
                if (time_kernels) W2B_CUDA_CHECK(cudaEventRecord(tev[4 * i + 3], 0));
                launches += 2;
#else
                for (int i2 = 0; i2 < SC_STRIDE; ++i2) C.sched[(parity ^ 1) * SC_STRIDE + i2] = 0;
                for (int c = 0; c < nc; ++c) { ProposeOut o{0, 0, -1, 0, -1, 0, 0, 0}; Ctx x; phase_propose(T, R, C, c, nsteps, parity, log_chain, dlog, &o, x); propose_flush(C, parity, &o, 1); }
                int* sc = C.sched + parity * SC_STRIDE;
                for (int c = 0; c < nc; ++c)
                    { Ctx x; phase_rare(T, R, C, A, c, nsteps, log_chain, dlog, ws.data(), (int*)(ws.data() + 3 * C.kcap + mat_doubles(C, true)), ws.data() + 3 * C.kcap, x); }
                if (C.dmax > 1) for (int t = 0; t < sc[SC_NT0] + sc[SC_NT1] + sc[SC_NT2]; ++t) phase_trace(T, C, t, parity, ws.data() + 3 * C.kcap);
                const int total = sc[SC_NRFIN] + sc[SC_NQ0] + sc[SC_NQ1] + sc[SC_NQ2];
                for (int t = 0; t < total; ++t)
                    { Ctx x; phase_decide(T, R, C, A, t, parity, nsteps, log_chain, dlog, ws.data(), (int*)(ws.data() + 3 * C.kcap + mat_doubles(C, true)), ws.data() + 3 * C.kcap, x); }
#endif
            }
            rounds += chunk;
            // chains that were runnable at the start of the last round, and the largest number of moves one of them had left
#if W2B_DEVICE
            W2B_CUDA_CHECK(cudaGetLastError());
            W2B_CUDA_CHECK(cudaMemcpyAsync(h_sched, C.sched + ((round - 1) & 1) * SC_STRIDE, SC_STRIDE * sizeof(int), cudaMemcpyDeviceToHost, 0));
            W2B_CUDA_CHECK(cudaStreamSynchronize(0));
            if (time_kernels) for (long long i = 0; i < chunk; ++i) {
                float t = 0;
                W2B_CUDA_CHECK(cudaEventElapsedTime(&t, tev[4 * i], tev[4 * i + 1])); ms_propose += t;
                W2B_CUDA_CHECK(cudaEventElapsedTime(&t, tev[4 * i + 1], tev[4 * i + 2])); ms_trace += t;
                W2B_CUDA_CHECK(cudaEventElapsedTime(&t, tev[4 * i + 2], tev[4 * i + 3])); ms_decide += t;
            }
#else
            std::memcpy(h_sched, C.sched + ((round - 1) & 1) * SC_STRIDE, SC_STRIDE * sizeof(int));
#endif
            if (h_sched[SC_NRUN] == 0) break;                 // nothing runnable and nothing in flight on the rare stream
            chunk = std::max<long long>(1, (long long)h_sched[SC_MAXLEFT] / prun);
        }
#if W2B_DEVICE
        W2B_CUDA_CHECK(cudaStreamSynchronize(rare_stream));       // the last rare move may still be storing its chain record
#endif
        if (log_chain >= 0) { sync(); d2h(log_host, dlog, (size_t)nsteps * LOGW); dev_free(dlog); }
    }

    void measure() {
#if W2B_DEVICE
        size_t sm = smem_bytes() + (size_t)WARPS_PER_CTA * T.nf * R.nlegmax * sizeof(double);
        if (sm > 48 * 1024 && sm > smem_cfg_measure) { W2B_CUDA_CHECK(cudaFuncSetAttribute(k_measure, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); smem_cfg_measure = sm; }
        int dev = 0, nsm = 1;
        W2B_CUDA_CHECK(cudaGetDevice(&dev));
        W2B_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
        const int per_bin = (C.nchains + R.nbins - 1) / R.nbins;                     // chains of one bin
        const int nsub = std::max(1, std::min((nsm * 8 + R.nbins - 1) / R.nbins, (per_bin + WARPS_PER_CTA - 1) / WARPS_PER_CTA));
        if (C.dmax > 1 && R.meas_dm) {
            if (!sw_ready) plan_sweep();
            k_meas_extend<<<tr_grid, TRACE_WARPS * 32, tr_smem>>>(T, C, tr_doubles); ++launches;
            C.meas_pre = 1;
        }
        k_measure<<<R.nbins * nsub, WARPS_PER_CTA * 32, sm>>>(T, R, C, A, wdoubles, mat_doubles(C, true), nsub); ++launches;
        C.meas_pre = 0;
        if (R.ngiw > 0) {
            const int per_cta = GIW_THREADS * GIW_F, ntile = (2 * R.ngiw + per_cta - 1) / per_cta;
            const int gsub = std::max(1, std::min(per_bin, (nsm * 16) / std::max(1, R.nbins * T.nf * ntile)));
            k_meas_giw<<<dim3(R.nbins * gsub, T.nf, ntile), GIW_THREADS>>>(T, R, C, A, gsub); ++launches;
        }
        W2B_CUDA_CHECK(cudaGetLastError());
#else
        std::vector<double> ws(wdoubles + 8);
        if (C.dmax > 1 && R.meas_dm) {
            for (int t = 0; t < C.nchains * 4; ++t) trace_job(T, C, (t >> 2) | ((t & 3) << Q_NTSHIFT) | Q_MEAS, ws.data());
            C.meas_pre = 1;
        }
        for (int c = 0; c < C.nchains; ++c) measure_chain(T, R, C, A, c, ws.data(), (int*)(ws.data() + 3 * C.kcap + mat_doubles(C, true)), ws.data() + 3 * C.kcap);
        C.meas_pre = 0;
        if (R.ngiw > 0) {                               // same per-thread routine as k_meas_giw, one "thread" per GIW_F frequencies
            const int nfreq = 2 * R.ngiw;
            for (int c = 0; c < C.nchains; ++c) for (int f = 0; f < T.nf; ++f) for (int m0 = 0; m0 < nfreq; m0 += GIW_F) {
                double re[GIW_F] = {0}, im[GIW_F] = {0};
                giw_accumulate_chain(T, R, C, c, f, m0, re, im);
                double* acc = A.base + (size_t)((R.chain_base + (unsigned)c) % (unsigned)R.nbins) * A.stride + A.o_giw + 2 * ((size_t)f * nfreq + m0);
                for (int q = 0; q < GIW_F && m0 + q < nfreq; ++q) { acc[2 * q] += re[q]; acc[2 * q + 1] += im[q]; }
            }
        }
#endif
    }

    void rebuild(int only_chain = -1) {
#if W2B_DEVICE
        size_t sm = smem_bytes();
        if (sm > 48 * 1024 && sm > smem_cfg_rebuild) { W2B_CUDA_CHECK(cudaFuncSetAttribute(k_rebuild, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); smem_cfg_rebuild = sm; }
        k_rebuild<<<grid(), WARPS_PER_CTA * 32, sm>>>(T, R, C, wdoubles, mat_doubles(C, true), only_chain); ++launches;
        W2B_CUDA_CHECK(cudaGetLastError());
#else
        std::vector<double> ws(wdoubles + 8);
        for (int c = 0; c < C.nchains; ++c) if (only_chain < 0 || c == only_chain) rebuild_chain(T, R, C, c, ws.data(), (int*)(ws.data() + 3 * C.kcap + mat_doubles(C, true)), ws.data() + 3 * C.kcap);
#endif
    }

    // ------------------------------------------------------------ host access to chain state
    void get_header(int chain, std::vector<int>& ih, std::vector<double>& dh) {
        ih.resize(IH_STRIDE); dh.resize(DH_STRIDE);
        sync();
        d2h(ih.data(), C.ih + (size_t)chain * IH_STRIDE, IH_STRIDE);
        d2h(dh.data(), C.dh + (size_t)chain * DH_STRIDE, DH_STRIDE);
    }
    void get_ops(int chain, std::vector<double>& tau, std::vector<unsigned char>& code, std::vector<short>& sst) {
        std::vector<int> ih; std::vector<double> dh; get_header(chain, ih, dh);
        int n = ih[IH_NOPS];
        tau.resize(n); code.resize(n); sst.resize(n);
        d2h(tau.data(), C.op_tau + (size_t)chain * C.ops_cap, n);
        d2h(code.data(), C.op_code + (size_t)chain * C.ops_cap, n);
        d2h(sst.data(), C.op_sst + (size_t)chain * C.ops_cap, n);
    }
    void get_minv(int chain, int f, std::vector<double>& out, int& k) {
        std::vector<int> ih; std::vector<double> dh; get_header(chain, ih, dh);
        k = ih[IH_KF + f];
        std::vector<double> buf((size_t)C.kcap * C.ld);
        d2h(buf.data(), C.M + ((size_t)chain * (T.nf + 1) + ih[IH_MSLOT + f]) * ((size_t)C.kcap * C.ld), buf.size());
        out.assign((size_t)k * k, 0.0);                // column-major (creator, annihilator) like the reference
        for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) out[i + (size_t)j * k] = buf[(size_t)i * C.ld + j];
    }
    void set_ops(int chain, int n, const double* tau, const unsigned char* code, int outer_sst, int outer_state = 0) {
        if (n > C.ops_cap) throw std::runtime_error("configuration exceeds operator capacity");
        sync();
        std::vector<int> ih(IH_STRIDE);
        d2h(ih.data(), C.ih + (size_t)chain * IH_STRIDE, IH_STRIDE);
        ih[IH_NOPS] = n; ih[IH_OUTER] = outer_sst; ih[IH_OSTATE] = outer_state; ih[IH_ERR] = 0;
        h2d(C.ih + (size_t)chain * IH_STRIDE, ih.data(), IH_STRIDE);
        h2d(C.op_tau + (size_t)chain * C.ops_cap, tau, n);
        h2d(C.op_code + (size_t)chain * C.ops_cap, code, n);
        rebuild(chain);
    }
    void get_counters(std::vector<unsigned long long>& out) {   // summed over chains
        sync();
        std::vector<unsigned long long> all((size_t)C.nchains * C_NCNT);
        d2h(all.data(), C.counters, all.size());
        out.assign(C_NCNT, 0ull);
        for (int c = 0; c < C.nchains; ++c) for (int i = 0; i < C_NCNT; ++i) out[i] += all[(size_t)c * C_NCNT + i];
    }
    int* d_flags = nullptr;
    int forced_err = 0;                        // error flags of other ranks (w2b_allreduce_accum)
    void force_error(int e) { forced_err |= e; }
    int* flag_scratch() { if (!d_flags) d_flags = keep(dev_alloc<int>(4)); return d_flags; }
    int error_flags(int* max_k = nullptr, int* max_ops = nullptr) {
#if W2B_DEVICE
        if (!d_flags) d_flags = keep(dev_alloc<int>(4));
        dev_zero(d_flags, 4);
        k_flags<<<(C.nchains + 255) / 256, 256>>>(C, T.nf, d_flags); ++launches;
        int h[4]; d2h(h, d_flags, 4);                         // synchronises
        if (max_k) *max_k = h[2];
        if (max_ops) *max_ops = h[1];
        return h[0] | forced_err;
#endif
        sync();
        std::vector<int> ih((size_t)C.nchains * IH_STRIDE);
        d2h(ih.data(), C.ih, ih.size());
        int e = 0, mk = 0, mo = 0;
        for (int c = 0; c < C.nchains; ++c) {
            e |= ih[(size_t)c * IH_STRIDE + IH_ERR];
            mo = std::max(mo, ih[(size_t)c * IH_STRIDE + IH_NOPS]);
            for (int f = 0; f < T.nf; ++f) mk = std::max(mk, ih[(size_t)c * IH_STRIDE + IH_KF + f]);
        }
        if (max_k) *max_k = mk;
        if (max_ops) *max_ops = mo;
        return e | forced_err;
    }
    double mean_order() {                       // <k_f> averaged over chains and flavours
        sync();
        std::vector<int> ih((size_t)C.nchains * IH_STRIDE);
        d2h(ih.data(), C.ih, ih.size());
        double s = 0; for (int c = 0; c < C.nchains; ++c) for (int f = 0; f < T.nf; ++f) s += ih[(size_t)c * IH_STRIDE + IH_KF + f];
        return s / ((double)C.nchains * T.nf);
    }
    void download_accum(std::vector<double>& out, std::vector<unsigned long long>& accpair) {
        sync();
        out.resize((size_t)R.nbins * accum_doubles);
        d2h(out.data(), A.base, out.size());
        accpair.resize((size_t)T.nf * R.naccpair);
        d2h(accpair.data(), A.accpair, accpair.size());
    }
};

}  // namespace w2b
