// Warp-cooperative primitives used by the chain engine (engine_core.hpp).
//
// On the GPU one warp owns one Markov chain: every routine is written as
// "for (i = lane; i < n; i += W2B_WARP)" loops plus the reductions below, with
// __syncwarp() ordering the chain's global-memory state between phases.
//
// The same source also compiles with W2B_HOST_EMU (g++, W2B_WARP == 1) into a
// test-only library so that the *logic* of the kernels can be checked against the
// oracle on machines without a GPU (tests/test_engine_logic.py).  That build is not
// reachable from the product API (w2dynamics_b200/CTQMC.py loads the CUDA library only
// and fails loudly without it).
#pragma once
#include <cstdint>

#if defined(__CUDACC__) && !defined(W2B_HOST_EMU)
#define W2B_DEVICE 1
#define W2B_HD __device__ __forceinline__
#define W2B_HDN __device__ __noinline__
#define W2B_WARP 32
#define W2B_FULL 0xffffffffu
#else
#define W2B_DEVICE 0
#define W2B_HD inline
#define W2B_HDN inline
#define W2B_WARP 1
#include <cmath>
#include <cstring>
#endif

namespace w2b {

#if W2B_DEVICE
W2B_HD int lane_id() { return threadIdx.x & 31; }
W2B_HD void wsync() { __syncwarp(); }
W2B_HD double wsum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(W2B_FULL, x, o);
    return x;
}
W2B_HD int wsum(int x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(W2B_FULL, x, o);
    return x;
}
W2B_HD double wmaxd(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(W2B_FULL, x, o));
    return x;
}
W2B_HD int wmax(int x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = max(x, __shfl_xor_sync(W2B_FULL, x, o));
    return x;
}
W2B_HD int wmin(int x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = min(x, __shfl_xor_sync(W2B_FULL, x, o));
    return x;
}
W2B_HD int wxor(int x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x ^= __shfl_xor_sync(W2B_FULL, x, o);
    return x;
}
W2B_HD bool wany(bool p) { return __any_sync(W2B_FULL, p); }
W2B_HD unsigned wballot(bool p) { return __ballot_sync(W2B_FULL, p); }
template <class T> W2B_HD T wbcast(T x, int src) { return __shfl_sync(W2B_FULL, x, src); }
template <class T> W2B_HD T wshfl_up1(T x) { return __shfl_up_sync(W2B_FULL, x, 1); }     // lane 0 keeps its own
template <class T> W2B_HD T wshfl_down1(T x) { return __shfl_down_sync(W2B_FULL, x, 1); } // last lane keeps its own
// inclusive prefix sum over lanes
W2B_HD int wscan_incl(int x) {
    int l = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(W2B_FULL, x, o); if (l >= o) x += y; }
    return x;
}
// arg-max of |x| over lanes: returns lane index holding the max (lowest lane on ties), value via reference
W2B_HD int wargmax_abs(double x, int idx, double& best) {
    double v = fabs(x);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(W2B_FULL, v, o);
        int oi = __shfl_xor_sync(W2B_FULL, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    best = v;
    return idx;
}
W2B_HD void atomic_add(double* p, double v) { atomicAdd(p, v); }
W2B_HD void atomic_add(unsigned long long* p, unsigned long long v) { atomicAdd(p, v); }
W2B_HD int atomic_add_int(int* p, int v) { return atomicAdd(p, v); }
// flags exchanged between kernels that run concurrently on different streams
W2B_HD int vload_u8(const unsigned char* p) { return (int)*(const volatile unsigned char*)p; }
W2B_HD void vstore_u8(unsigned char* p, unsigned char v) { *(volatile unsigned char*)p = v; }
W2B_HD void fence_device() { __threadfence(); }
W2B_HD void atomic_max_int(int* p, int v) { atomicMax(p, v); }
// CTA-level primitives of the phase scheduler
W2B_HD void cta_sync() { __syncthreads(); }
W2B_HD int cta_warp() { return threadIdx.x >> 5; }
W2B_HD int cta_nwarps() { return blockDim.x >> 5; }
W2B_HD int cta_tid() { return threadIdx.x; }
W2B_HD int cta_nthreads() { return blockDim.x; }
// lane 0 takes the next ticket of a shared-memory counter, all lanes get it
W2B_HD int wgrab(int* ctr) { int v = 0; if ((threadIdx.x & 31) == 0) v = atomicAdd(ctr, 1); return __shfl_sync(W2B_FULL, v, 0); }
W2B_HD int wgrab_n(int* ctr, int n) { int v = 0; if ((threadIdx.x & 31) == 0) v = atomicAdd(ctr, n); return __shfl_sync(W2B_FULL, v, 0); }
// single-thread ticket (caller is one lane)
W2B_HD int wgrab_one(int* ctr) { return atomicAdd(ctr, 1); }
W2B_HD double ldg(const double* p) { return __ldg(p); }
W2B_HD int ldg(const int* p) { return __ldg(p); }
W2B_HD short ldg(const short* p) { return __ldg(p); }
W2B_HD long long ldg(const long long* p) { return __ldg(p); }
#else
W2B_HD int lane_id() { return 0; }
W2B_HD void wsync() {}
W2B_HD double wsum(double x) { return x; }
W2B_HD int wsum(int x) { return x; }
W2B_HD double wmaxd(double x) { return x; }
W2B_HD int wmax(int x) { return x; }
W2B_HD int wmin(int x) { return x; }
W2B_HD int wxor(int x) { return x; }
W2B_HD bool wany(bool p) { return p; }
W2B_HD unsigned wballot(bool p) { return p ? 1u : 0u; }
template <class T> W2B_HD T wbcast(T x, int) { return x; }
template <class T> W2B_HD T wshfl_up1(T x) { return x; }
template <class T> W2B_HD T wshfl_down1(T x) { return x; }
W2B_HD int wscan_incl(int x) { return x; }
W2B_HD int wargmax_abs(double x, int idx, double& best) { best = std::fabs(x); return idx; }
W2B_HD void atomic_add(double* p, double v) { *p += v; }
W2B_HD void atomic_add(unsigned long long* p, unsigned long long v) { *p += v; }
W2B_HD int atomic_add_int(int* p, int v) { int o = *p; *p += v; return o; }
W2B_HD int vload_u8(const unsigned char* p) { return (int)*p; }
W2B_HD void vstore_u8(unsigned char* p, unsigned char v) { *p = v; }
W2B_HD void fence_device() {}
W2B_HD void atomic_max_int(int* p, int v) { if (v > *p) *p = v; }
W2B_HD void cta_sync() {}
W2B_HD int cta_warp() { return 0; }
W2B_HD int cta_nwarps() { return 1; }
W2B_HD int cta_tid() { return 0; }
W2B_HD int cta_nthreads() { return 1; }
W2B_HD int wgrab(int* ctr) { return (*ctr)++; }
W2B_HD int wgrab_n(int* ctr, int n) { int v = *ctr; *ctr += n; return v; }
W2B_HD int wgrab_one(int* ctr) { return (*ctr)++; }
W2B_HD double ldg(const double* p) { return *p; }
W2B_HD int ldg(const int* p) { return *p; }
W2B_HD short ldg(const short* p) { return *p; }
W2B_HD long long ldg(const long long* p) { return *p; }
#endif

// Out-of-line math: log / exp expand to ~50-100 SASS instructions each; one shared copy keeps the hot path inside the
// 32 KB L1.5 instruction cache (warps of different chains sit at different program counters).
W2B_HDN double w_log(double x) { return log(x); }
W2B_HDN double w_exp(double x) { return exp(x); }
W2B_HDN double wsumd(double x) { return wsum(x); }
// sin(pi t), cos(pi t) with exact argument reduction
#if W2B_DEVICE
W2B_HD void w_sincospi(double t, double& s, double& c) { sincospi(t, &s, &c); }
#else
W2B_HD void w_sincospi(double t, double& s, double& c) { const double r = t - 2.0 * std::floor(0.5 * t); s = std::sin(M_PI * r); c = std::cos(M_PI * r); }
#endif

// Unfused multiply / add: proposal times must be bit-identical to a CPU evaluation of the reference formulas
// (nvcc would contract a*b+c into one FMA; times feed exact comparisons and integer bookkeeping).
#if W2B_DEVICE
W2B_HD double mul_rn(double a, double b) { return __dmul_rn(a, b); }
W2B_HD double add_rn(double a, double b) { return __dadd_rn(a, b); }
#else
W2B_HD double mul_rn(double a, double b) { return a * b; }
W2B_HD double add_rn(double a, double b) { return a + b; }
#endif

// ---------------------------------------------------------------- counter-based RNG
// Philox4x32-10 keyed by (seed, global chain id); draw n uses words (2*(n&1), 2*(n&1)+1) of block n>>1.
// Every lane evaluates the same stream (uniform control flow), so no shuffles are needed.
struct Philox {
    uint32_t k0, k1;
    uint64_t ctr;
    uint32_t buf[4];
    uint64_t have;
    W2B_HD void init(uint32_t seed, uint32_t chain, uint64_t counter) { k0 = seed; k1 = chain; ctr = counter; have = ~0ull; }
    W2B_HDN void block(uint64_t blk) {
        uint32_t c0 = (uint32_t)blk, c1 = (uint32_t)(blk >> 32), c2 = 0, c3 = 0, a = k0, b = k1;
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
            uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ a, n1 = (uint32_t)p1;
            uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ b, n3 = (uint32_t)p0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3; a += W0; b += W1;
        }
        buf[0] = c0; buf[1] = c1; buf[2] = c2; buf[3] = c3; have = blk;
    }
    W2B_HD double next() {
        uint64_t blk = ctr >> 1;
        if (have != blk) block(blk);
        uint32_t x = (ctr & 1) ? buf[2] : buf[0], y = (ctr & 1) ? buf[3] : buf[1];
        ++ctr;
        return ((double)(x >> 5) * 67108864.0 + (double)(y >> 6)) / 9007199254740992.0;
    }
    W2B_HD void skip(int n) { ctr += (uint64_t)n; }       // draws whose value is never used (counter-based: free)
    W2B_HD int randint(int lo, int hi) {
        int r = lo + (int)(next() * (double)(hi - lo + 1));
        return r > hi ? hi : r;
    }
};

}  // namespace w2b
