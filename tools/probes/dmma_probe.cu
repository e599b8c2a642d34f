// Micro-benchmarks behind DESIGN.md's FP64 roofline discussion (run on the B200 box: tools/probes/run_probes.sh).
//   * DMMA.8x8x4 issue rate per SM with 1..8 warps per scheduler and 1..8 independent accumulator chains per warp
//     (sm_100a lowers every mma.sync f64 shape to DMMA.8x8x4 -- checked with cuobjdump);
//   * dependent-chain latency of DMMA.8x8x4;
//   * cost of exp(): CUDA libdevice exp() vs. the table + degree-5 polynomial used by the trace kernel.
#include <cstdio>
#include <cuda_runtime.h>
#include <cmath>
#include <vector>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int CH> __global__ void k_dmma(double* out, int iters, double a0, double b0) {
    double acc[CH][2];
#pragma unroll
    for (int c = 0; c < CH; ++c) { acc[c][0] = 0.0; acc[c][1] = 0.0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) dmma(acc[c][0], acc[c][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += acc[c][0] + acc[c][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma(double* out, int iters, double a0, double b0) {
    double acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = fma(acc[c], a0, b0);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ double fast_exp_neg(double x, const double* tab) {
    // x <= 0: exp(x) = 2^(k/64) * exp(r), k = rint(x * 64 / ln2)
    const double t = x * 92.332482616893656877;                      // 64 / ln 2
    const double kf = t + 6755399441055744.0;                         // 1.5 * 2^52: round to nearest integer
    const int k = __double2loint(kf);
    const double kd = kf - 6755399441055744.0;
    double r = fma(kd, -1.08304246962326283e-02, x);                  // ln2/64 hi
    r = fma(kd, -4.0e-19 * 0.0, r);
    double p = 1.0 / 120.0;
    p = fma(p, r, 1.0 / 24.0); p = fma(p, r, 1.0 / 6.0); p = fma(p, r, 0.5); p = fma(p, r, 1.0); p = fma(p, r, 1.0);
    const double tv = tab[k & 63];
    const int e = k >> 6;
    double res = tv * p;
    res = __hiloint2double(__double2hiint(res) + (e << 20), __double2loint(res));
    return e < -1000 ? 0.0 : res;
}
__global__ void k_exp(double* out, int iters, double x0, int mode) {
    __shared__ double tab[64];
    if (threadIdx.x < 64) tab[threadIdx.x] = exp2((double)threadIdx.x / 64.0);
    __syncthreads();
    double s = 0.0, x = -x0 * (1 + threadIdx.x);
    for (int i = 0; i < iters; ++i) { s += mode ? fast_exp_neg(x, tab) : exp(x); x -= 1e-3; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double* out; cudaMalloc(&out, 148 * 1024 * 8 * 4);
    const int iters = 20000;
    printf("{\"nsm\": %d, \"clock_khz\": %d, \"dmma\": [", nsm, khz);
    bool first = true;
#define RUN(CH) for (int wps = 1; wps <= 8; wps *= 2) { \
        float ms = timeit([&] { k_dmma<CH><<<nsm, wps * 128>>>(out, iters, 1.0, 1e-3); }); \
        double n = (double)nsm * wps * 4 * iters * CH; \
        printf("%s{\"chains\": %d, \"warps_per_sched\": %d, \"ms\": %.4f, \"tflops\": %.3f, \"ns_per_dmma_per_sched\": %.3f}", first ? "" : ", ", CH, wps, ms, n * 512 / ms / 1e9, ms * 1e6 / (iters * CH * wps)); first = false; }
    RUN(1) RUN(2) RUN(4) RUN(8)
    printf("], ");
    { float ms = timeit([&] { k_dfma<<<nsm, 1024>>>(out, iters, 1.0000001, 1e-3); });
      printf("\"dfma_tflops\": %.3f, ", (double)nsm * 1024 * iters * 8 * 2 / ms / 1e9); }
    for (int mode = 0; mode < 2; ++mode) {
        float ms = timeit([&] { k_exp<<<nsm, 1024>>>(out, 2000, 1e-2, mode); });
        printf("\"exp_%s_ns_per_warp_call_per_sched\": %.3f, ", mode ? "table" : "libdevice", ms * 1e6 / (2000.0 * 8));
    }
    printf("\"note\": \"ns_per_dmma_per_sched x clock = cycles one scheduler needs per DMMA.8x8x4\"}\n");
    return 0;
}
