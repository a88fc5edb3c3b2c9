// k6_thin_cluster.cu -- the thread-block-cluster kernel of the thin solver (per-instance factors, n >= 512: C5)
#include "gi_thin.cuh"

namespace cb {

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


cudaError_t gt_launch_cluster(const GtBatch& B, const GtPlan& plan, cudaStream_t st) { return gt_launch_cluster_t<512>(B, plan, st); }

} // namespace cb
