// fp64_peak.cu -- measured FP64 ceilings of the device the engine runs on (SURVEY.md 8d asks the roofline to be
// stated against MEASURED DFMA and DMMA.8x8x4 throughput, not the nominal data-sheet figure).
//   dfma : 8 independent FMA chains per thread, 2 flop per FMA
//   dmma : 8 independent mma.sync.m8n8k4.f64 accumulator tiles per warp, 512 flop per instruction
#include "launch.h"

#include <cuda_runtime.h>

namespace cb {

namespace {

constexpr int kChains = 8;
constexpr int kInnerFma = 4096; // ~0.6 ms per launch
constexpr int kInnerMma = 1024; // ~1.1 ms per launch

__global__ void __launch_bounds__(256) dfma_chain_kernel(double* out, double a, double b)
{
    double acc[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) acc[i] = double(threadIdx.x + i);
    for (int it = 0; it < kInnerFma; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += acc[i];
    if (s == 123.456) out[0] = s; // never true for the arguments used; keeps the chains alive
}

__global__ void __launch_bounds__(256) dmma_chain_kernel(double* out, double a, double b)
{
    double c0[kChains], c1[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) { c0[i] = double(threadIdx.x); c1[i] = double(i); }
    for (int it = 0; it < kInnerMma; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[i]), "+d"(c1[i])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}

} // namespace

// returns 0 or a negative cudaError_t; TFLOP/s, best of 3 timed launches after one warm-up
int fp64_peaks_measure(int sms, double* scratch, cudaStream_t st, double* dfma_tflops, double* dmma_tflops)
{
    cudaEvent_t e0, e1;
    cudaError_t e;
    if ((e = cudaEventCreate(&e0)) != cudaSuccess) return -int(e);
    if ((e = cudaEventCreate(&e1)) != cudaSuccess) { cudaEventDestroy(e0); return -int(e); }
    const int grid = sms * 8, threads = 256;
    double best[2] = { 0.0, 0.0 };
    for (int which = 0; which < 2; ++which) {
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0, st);
            if (which == 0) dfma_chain_kernel<<<grid, threads, 0, st>>>(scratch, 0.999999, 1e-9);
            else dmma_chain_kernel<<<grid, threads, 0, st>>>(scratch, 0.5, 1e-3);
            cudaEventRecord(e1, st);
            if ((e = cudaEventSynchronize(e1)) != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); return -int(e); }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            const double per_thread = double(which == 0 ? kInnerFma : kInnerMma) * kChains * (which == 0 ? 2.0 : 512.0 / 32.0);
            const double tf = per_thread * double(grid) * threads / (double(ms) * 1e-3) * 1e-12;
            if (rep > 0 && tf > best[which]) best[which] = tf;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if ((e = cudaGetLastError()) != cudaSuccess) return -int(e);
    *dfma_tflops = best[0];
    *dmma_tflops = best[1];
    return 0;
}

} // namespace cb
