// common.cuh -- shared helpers for the copra_b200 CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace cb {

constexpr int kWarp = 32;
constexpr int kMaxWarps = 32;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// (value, index) arg-min with strict '<' on value and lowest index on ties (qpgen2's forward scans).
struct MinIdx {
    double v;
    int i;
};
__device__ __forceinline__ MinIdx better(MinIdx a, MinIdx b)
{
    if (b.i >= 0 && (a.i < 0 || b.v < a.v || (b.v == a.v && b.i < a.i))) return b;
    return a;
}
// Warp arg-min through the redux unit: the doubles are mapped to order-preserving 64-bit keys, the minimum key is
// found with two 32-bit redux.min (high word, then low word among the lanes that tie on it) and the lowest index
// among the lanes that hold it with a third -- same result as the pairwise `better` tournament (-0.0 is folded
// into +0.0 first so that the key order agrees with IEEE '<' / '==').
__device__ __forceinline__ MinIdx warp_argmin(MinIdx m)
{
    const bool has = m.i >= 0;
    unsigned long long u = (unsigned long long)__double_as_longlong(m.v + 0.0);
    u = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
    const unsigned hi = has ? unsigned(u >> 32) : 0xffffffffu, lo = unsigned(u);
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const bool c1 = has && hi == mh;
    const unsigned ml = __reduce_min_sync(0xffffffffu, c1 ? lo : 0xffffffffu);
    const unsigned idx = __reduce_min_sync(0xffffffffu, (c1 && lo == ml) ? unsigned(m.i) : 0xffffffffu);
    MinIdx r;
    if (idx == 0xffffffffu) { r.v = 0.0; r.i = -1; return r; }
    unsigned long long k = ((unsigned long long)mh << 32) | ml;
    k = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    r.v = __longlong_as_double((long long)k);
    r.i = int(idx);
    return r;
}

// Slack of bound row j in [0, 2n): upper rows first (-x_v >= -ub_v), then lower rows (x_v >= lb_v); `active` is indexed
// by constraint (general rows first, then the 2n bound rows).  A variable PINNED by lb == ub whose one bound is already
// active satisfies the other by definition: its slack is 0, not the rounding residue of x_v.  (A residue of -1e-16 would
// make the solver try to add the negated normal of an active row and report a spurious "infeasible"; this is the default
// state of InitialStateLMPC, whose x0 bounds are both ps->x0 until resetInitialStateBounds -- src/InitialStateLMPC.cpp:24-25.)
__device__ __forceinline__ double gi_bound_slack(int j, int n, int mg, const double* x, const double* lb, const double* ub,
    const unsigned char* active)
{
    const int v = j < n ? j : j - n;
    const double s = (j < n) ? ub[v] - x[v] : x[v] - lb[v];
    const int twin = mg + (j < n ? j + n : j - n);
    return (active[twin] && lb[v] == ub[v]) ? 0.0 : s;
}

// Block-wide reductions through a small smem scratch (>= kMaxWarps entries each).  All threads
// must call; result is returned to every thread.  Ends with a __syncthreads so scratch is reusable.
__device__ __forceinline__ double block_sum(double v, double* scratch)
{
    v = warp_sum(v);
    const int nw = (blockDim.x + 31) >> 5;
    if (lane_id() == 0) scratch[warp_id()] = v;
    __syncthreads();
    double t = (lane_id() < nw) ? scratch[lane_id()] : 0.0;
    t = warp_sum(t);
    __syncthreads();
    return t;
}
__device__ __forceinline__ void block_sum2(double& a, double& b, double* scratch)
{
    a = warp_sum(a);
    b = warp_sum(b);
    const int nw = (blockDim.x + 31) >> 5;
    if (lane_id() == 0) {
        scratch[warp_id()] = a;
        scratch[kMaxWarps + warp_id()] = b;
    }
    __syncthreads();
    double ta = (lane_id() < nw) ? scratch[lane_id()] : 0.0;
    double tb = (lane_id() < nw) ? scratch[kMaxWarps + lane_id()] : 0.0;
    a = warp_sum(ta);
    b = warp_sum(tb);
    __syncthreads();
}
__device__ __forceinline__ MinIdx block_argmin(MinIdx m, double* scratch_v, int* scratch_i)
{
    m = warp_argmin(m);
    const int nw = (blockDim.x + 31) >> 5;
    if (lane_id() == 0) {
        scratch_v[warp_id()] = m.v;
        scratch_i[warp_id()] = m.i;
    }
    __syncthreads();
    MinIdx t;
    t.v = (lane_id() < nw) ? scratch_v[lane_id()] : 0.0;
    t.i = (lane_id() < nw) ? scratch_i[lane_id()] : -1;
    t = warp_argmin(t);
    __syncthreads();
    return t;
}

} // namespace cb
