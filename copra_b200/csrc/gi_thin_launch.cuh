// gi_thin_launch.cuh -- the persistent one-CTA-per-instance kernel of the thin solver and its launcher, instantiated once per
// form by k6_thin_f0.cu / k6_thin_f1.cu / k6_thin_f2.cu (see gt_solve's FORM)
#pragma once
#include "gi_thin.cuh"

namespace cb {

template <int MAXT, int MINB, int FORM>
__global__ void __launch_bounds__(MAXT, MINB) gi_thin_kernel(const __grid_constant__ GtBatch B)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_next;
    GtWork W = gt_carve(B.lay, smem, B.ws + (long long)blockIdx.x * B.ws_stride, B.n);
    for (;;) {
        if (threadIdx.x == 0) s_next = atomicAdd(B.counter, 1);
        __syncthreads();
        const int q = s_next;
        __syncthreads();
        if (q >= B.batch) break;
        const int b = B.order ? B.order[q] : q; // longest-first when a prepass ranked the instances
        gt_solve<FORM>(GtSolo(), B, W, b, B.vsmall, B.max_iter);
        __syncthreads();
    }
}

template <int MAXT, int MINB, int FORM> static cudaError_t gt_launch_t(const GtBatch& B, const GtPlan& plan, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(gi_thin_kernel<MAXT, MINB, FORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(plan.smem_bytes));
    if (e != cudaSuccess) return e;
    gi_thin_kernel<MAXT, MINB, FORM><<<plan.grid, plan.threads, plan.smem_bytes, st>>>(B);
    return cudaGetLastError();
}


#define GT_DEFINE_FORM_LAUNCH(FORM)                                                                                        \
    cudaError_t gt_launch_form##FORM(const GtBatch& B, const GtPlan& plan, cudaStream_t st)                                \
    {                                                                                                                      \
        return plan.per_sm >= 2 ? gt_launch_t<512, 2, FORM>(B, plan, st) : gt_launch_t<512, 1, FORM>(B, plan, st);           \
    }

} // namespace cb
