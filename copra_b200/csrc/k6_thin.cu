// k6_thin.cu -- kernels of the throughput-oriented solver (gi_thin.cuh): the once-per-Hessian factorisation and the
// persistent one-CTA-per-instance active-set kernel.
#include "gi_thin.cuh"
#include "launch.h"

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstdlib>

namespace cb {

// Jt = R^-1 (Q = R'R) for `count` Hessians, one CTA each (blocked DMMA Cholesky + inverse of gi_factor.cuh working in
// global memory / L2), plus the transposed copy the row-oriented mat-vec reads.  pd[h] = 0 when Q is not positive
// definite (QuadProg's fail code 2).
__global__ void __launch_bounds__(512, 1) gt_factor_kernel(DArr Q, int n, int ld, int count, double* __restrict__ Jt,
    double* __restrict__ JtT, int* __restrict__ pd)
{
    extern __shared__ __align__(16) unsigned char smem[];
    double* scratch = reinterpret_cast<double*>(smem);
    const int tid = threadIdx.x, T = blockDim.x;
    const size_t sz = size_t(ld) * n;
    for (int h = blockIdx.x; h < count; h += gridDim.x) {
        double* J = Jt + (long long)h * sz;
        double* JT = JtT + (long long)h * sz;
        const double* q = Q.at(h);
        __syncthreads();
        for (size_t idx = tid; idx < sz; idx += T) {
            const int i = int(idx % ld), j = int(idx / ld);
            J[idx] = (i < n) ? q[i + size_t(j) * n] : 0.0;
            JT[idx] = 0.0;
        }
        __syncthreads();
        const bool ok = gi_factor_blocked(J, ld, n, scratch);
        __syncthreads();
        if (tid == 0) pd[h] = ok ? 1 : 0;
        if (!ok) continue;
        // transpose through 32 x 32 tiles (scratch holds >= 33 * 32 doubles)
        double* tile = scratch;
        const int nt = (n + 31) >> 5;
        for (int t = 0; t < nt * nt; ++t) {
            const int ti = t % nt, tj = t / nt; // tile rows ti*32.., cols tj*32..
            if (ti > tj) continue;              // strictly lower tiles are zero
            __syncthreads();
            for (int e = tid; e < 1024; e += T) {
                const int i = ti * 32 + (e & 31), j = tj * 32 + (e >> 5);
                tile[(e >> 5) * 33 + (e & 31)] = (i < n && j < n) ? J[i + size_t(j) * ld] : 0.0;
            }
            __syncthreads();
            for (int e = tid; e < 1024; e += T) {
                const int j = tj * 32 + (e & 31), i = ti * 32 + (e >> 5);
                if (i < n && j < n) JT[j + size_t(i) * ld] = tile[(e & 31) * 33 + (e >> 5)];
            }
        }
    }
}

// Psi' (nU x X, column-major) of ONE system from its compact first block column Gs: Psi[s nx + a, j nu + b] = (A^(s-1-j) B)[a, b]
// for j < s, zero otherwise (src/PreviewSystem.cpp:63-71).  Right-hand operand of Dpsi = Jt' Psi'.
__global__ void gt_psit_fill_kernel(const double* __restrict__ Gs, double* __restrict__ PsiT, int nx, int nu, int N)
{
    const int nU = nu * N, X = nx * (N + 1);
    const long long total = (long long)nU * X;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int i = int(idx % nU), c = int(idx / nU);
        const int j = i / nu, bb = i - j * nu, st = c / nx, a = c - st * nx;
        PsiT[idx] = (j < st) ? Gs[size_t(st - 1 - j) * nx + a + size_t(bb) * N * nx] : 0.0;
    }
}

cudaError_t gt_psit_fill_launch(const double* Gs, double* PsiT, int nx, int nu, int N, cudaStream_t st)
{
    gt_psit_fill_kernel<<<296, 256, 0, st>>>(Gs, PsiT, nx, nu, N);
    return cudaGetLastError();
}

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

// One thread-block CLUSTER per instance (GtClus): the CTAs split every stream, keep identical replicas of the small vectors in
// their shared memories (remote stores over DSMEM) and pull instances from the same queue (rank 0 pops, the index is stored
// into every CTA's s_next).
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) gi_thin_cluster_kernel(const __grid_constant__ GtBatch B)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_next;
    cg::cluster_group cgc = cg::this_cluster();
    GtClus cl;
    cl.r = int(cgc.block_rank()); cl.c = int(cgc.num_blocks());
    // cooperative phase: the first `kheavy` entries of the (longest-first) queue, one cluster per instance, on the workspace
    // of the cluster's first CTA
    GtWork W = gt_carve(B.lay, smem, B.ws + (long long)(blockIdx.x - cl.r) * B.ws_stride, B.n);
    int q;
    for (;;) {
        if (cl.r == 0 && threadIdx.x == 0) {
            const int v = atomicAdd(B.counter, 1);
            for (int k = 0; k < cl.c; ++k) *cgc.map_shared_rank(&s_next, k) = v;
        }
        cl.sync();
        q = s_next;
        if (q >= B.batch) return;
        if (q >= B.kheavy) break;
        const int b = B.order ? B.order[q] : q;
        gt_solve<0>(cl, B, W, b, B.vsmall, B.max_iter);
        cl.sync(); // every CTA is done with this instance (and has read s_next) before the next index or a remote store arrives
    }
    // throughput phase: the rest of the queue, one CTA per instance (2.2x more work per SM-second than a cluster); the index
    // popped last goes to the first CTA.  No cluster-scope operation from here on.
    W = gt_carve(B.lay, smem, B.ws + (long long)blockIdx.x * B.ws_stride, B.n);
    if (cl.r != 0) q = -1;
    for (;;) {
        if (q < 0) {
            __syncthreads();
            if (threadIdx.x == 0) s_next = atomicAdd(B.counter, 1);
            __syncthreads();
            q = s_next;
        }
        if (q >= B.batch) break;
        const int b = B.order ? B.order[q] : q;
        gt_solve<0>(GtSolo(), B, W, b, B.vsmall, B.max_iter);
        q = -1;
    }
}

static int gt_env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

int gt_cluster_capacity(const GtPlan& plan, int csize);

GtPlan gt_plan(const GtShape& sh, int batch, int sms, size_t smem_optin)
{
    GtPlan p{};
    p.threads = gt_env_int("COPRA_B200_THIN_THREADS", 512);
    // 512 threads per instance: 256 (47.7 k solves/s on C3) and 1024 x 1 CTA/SM (51.1 k) were measured against 512 x 2 (63.0 k) and
    // dropped -- every extra instantiation of the solver costs minutes of compile time
    p.threads = 512;
    const size_t base = gt_layout(sh.n, sh.meq, sh.m, sh.tab_doubles, p.threads, 0, sh.ss_doubles).bytes;
    p.ok = base + 2048 <= smem_optin;
    if (!p.ok) return p;
    if (sh.cluster > 1) {
        // one cluster per instance: one CTA per SM, Q1 entirely in the global workspace (every CTA of the cluster reads it)
        p.cluster = sh.cluster;
        p.q1s = 0;
        p.per_sm = 1;
        p.smem_bytes = base;
        p.ws_stride = (long long)sh.ld * sh.n + (long long)sh.n * sh.n;
        const int cap = gt_cluster_capacity(p, sh.cluster);
        if (cap <= 0) { p.ok = 0; return p; }
        p.grid = std::max(1, std::min(batch, cap)) * sh.cluster;
        return p;
    }
    // one CTA per SM; whatever shared memory the vectors and tables leave holds the head columns of Q1
    int per_sm = std::max(1, gt_env_int("COPRA_B200_THIN_CTAS_PER_SM", 2));
    per_sm = std::min(per_sm, 2048 / p.threads);
    per_sm = std::min<int>(per_sm, int((smem_optin + 1024) / (base + 1024)));
    const size_t budget = (smem_optin + 1024) / per_sm - 1024 - 1024; // per CTA, minus static shared memory headroom
    int q1s = budget > base ? int((budget - base) / (size_t(sh.ld) * sizeof(double))) : 0;
    q1s = std::min(q1s, sh.n);
    q1s = std::min(q1s, std::max(0, gt_env_int("COPRA_B200_THIN_Q1S", sh.n)));
    p.q1s = q1s;
    p.smem_bytes = gt_layout(sh.n, sh.meq, sh.m, sh.tab_doubles, p.threads, q1s, sh.ss_doubles).bytes;
    p.per_sm = per_sm;
    p.grid = std::max(1, std::min(batch, sms * per_sm));
    p.grid = std::max(1, std::min(p.grid, gt_env_int("COPRA_B200_THIN_GRID", p.grid)));
    p.ws_stride = (long long)sh.ld * sh.n + (long long)sh.n * sh.n; // Q1 (or P), S
    return p;
}

size_t gt_factor_smem(int n) { return std::max<size_t>(gi_factor_scratch(n), 33 * 32) * sizeof(double); }

cudaError_t gt_factor_launch(DArr Q, int n, int ld, int count, double* Jt, double* JtT, int* pd, int sms, cudaStream_t st)
{
    const size_t smem = gt_factor_smem(n);
    cudaError_t e = cudaFuncSetAttribute(gt_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    // (two CTAs per SM fit for n = 800 but cap the kernel at 64 registers: 165 ms instead of 145 ms for 1024 Hessians)
    gt_factor_kernel<<<std::max(1, std::min(count, sms)), 512, smem, st>>>(Q, n, ld, count, Jt, JtT, pd);
    return cudaGetLastError();
}

template <int MAXT, int MINB, int FORM> static cudaError_t gt_launch_t(const GtBatch& B, const GtPlan& plan, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(gi_thin_kernel<MAXT, MINB, FORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(plan.smem_bytes));
    if (e != cudaSuccess) return e;
    gi_thin_kernel<MAXT, MINB, FORM><<<plan.grid, plan.threads, plan.smem_bytes, st>>>(B);
    return cudaGetLastError();
}

template <int MAXT> static void gt_cluster_config(const GtPlan& plan, int csize, int grid, cudaStream_t st, cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr)
{
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3(unsigned(grid));
    cfg.blockDim = dim3(unsigned(plan.threads));
    cfg.dynamicSmemBytes = plan.smem_bytes;
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = unsigned(csize);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
}

template <int MAXT> static int gt_cluster_capacity_t(const GtPlan& plan, int csize)
{
    if (cudaFuncSetAttribute(gi_thin_cluster_kernel<MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(plan.smem_bytes)) != cudaSuccess) return 0;
    if (csize > 8 && cudaFuncSetAttribute(gi_thin_cluster_kernel<MAXT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1];
    gt_cluster_config<MAXT>(plan, csize, csize, nullptr, cfg, attr);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gi_thin_cluster_kernel<MAXT>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// resident clusters of `csize` CTAs for this plan (0: cannot be scheduled)
int gt_cluster_capacity(const GtPlan& plan, int csize)
{
    return gt_cluster_capacity_t<512>(plan, csize);
}

template <int MAXT> static cudaError_t gt_launch_cluster_t(const GtBatch& B, const GtPlan& plan, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(gi_thin_cluster_kernel<MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(plan.smem_bytes));
    if (e != cudaSuccess) return e;
    if (plan.cluster > 8) {
        e = cudaFuncSetAttribute(gi_thin_cluster_kernel<MAXT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1];
    gt_cluster_config<MAXT>(plan, plan.cluster, plan.grid, st, cfg, attr);
    return cudaLaunchKernelEx(&cfg, gi_thin_cluster_kernel<MAXT>, B);
}

size_t gt_sort_temp_bytes(int count)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, (const int*)nullptr, (int*)nullptr, (const int*)nullptr, (int*)nullptr, count);
    return bytes;
}

cudaError_t gt_sort_launch(const int* keys, int* keys_sorted, const int* idx, int* order, int count, void* temp, size_t temp_bytes, cudaStream_t st)
{
    return cub::DeviceRadixSort::SortPairsDescending(temp, temp_bytes, keys, keys_sorted, idx, order, count, 0, 32, st);
}

cudaError_t gt_launch(const GtBatch& B, const GtPlan& plan, cudaStream_t st)
{
    if (plan.cluster > 1) return gt_launch_cluster_t<512>(B, plan, st);
    // shared-factor form (C3-like batches: 2 CTAs/SM whenever it applies) / general form
    if (B.Hpsi && B.ss && B.structured && !B.warm && !B.prekey)
        return plan.per_sm >= 2 ? gt_launch_t<512, 2, 2>(B, plan, st) : gt_launch_t<512, 1, 2>(B, plan, st);
    if (B.Hpsi) return plan.per_sm >= 2 ? gt_launch_t<512, 2, 1>(B, plan, st) : gt_launch_t<512, 1, 1>(B, plan, st);
    return plan.per_sm >= 2 ? gt_launch_t<512, 2, 0>(B, plan, st) : gt_launch_t<512, 1, 0>(B, plan, st);
}

} // namespace cb
