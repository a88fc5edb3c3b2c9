// launch.h -- host-callable launchers of the sm_100a kernels (internal to libcopra_b200.so).
#pragma once
#include "engine.cuh"
#include <cuda_runtime.h>

namespace cb {

struct GiBatch {
    int n, meq, m, batch;
    DArr Q, c, Aeq, beq, Aineq, bineq, lb, ub;
    double* x;    // n per instance (may be null)
    int* status;  // 1
    int* iters;   // 2
    int* nact;    // 1
    int* iact;    // n
    double* ws;   // global J / S workspace, ws_stride doubles per CTA (may be null if unused)
    long long ws_stride;
    int* counter; // zero-initialised work-queue cursor
    double vsmall;
    int max_iter;
    int j_smem, s_smem, a_smem;
    // gi_small_kernel only -- the factor J = R^-1 of every instance kept across a receding-horizon re-solve (the Hessian does
    // not change with x0): jmode 1 = store J (and jflag = 1 when Q was positive definite) after factoring, 2 = load it instead
    // of factoring (instances whose jflag is 0 factor again), 0 = off.  jcache: ld_vec2(n) * even(n) doubles per instance.
    double* jcache;
    int* jflag;
    int jmode;
};

struct GiPlan {
    int threads, grid, j_smem, s_smem, a_smem;
    int small; // 1: gi_small_kernel (n <= 64, all state in shared memory)
    int cluster; // > 0: gi_cluster_kernel with this cluster size (J distributed over DSMEM)
    size_t smem_bytes;
    long long ws_stride; // doubles per CTA of global workspace
};

// choose threads / smem residency / grid for a (n, meq, m) shape on a device with `sms` SMs
GiPlan gi_plan(int n, int meq, int m, int batch, int sms, size_t smem_optin);
cudaError_t gi_launch(const GiBatch& B, const GiPlan& plan, cudaStream_t st);
double gi_vsmall();
// cluster path (gi_cluster.cuh): S workspace of odd_ld(n)*n doubles per cluster, sized by the caller
int gi_cluster_max_clusters(const GiPlan& plan);
cudaError_t gi_cluster_launch(const GiBatch& B, const GiPlan& plan, double* Sws, int nclusters, cudaStream_t st);

// K1 .. K7 of the batched LMPC engine.  Each returns the number of kernels it launched (>0) or a
// negative cudaError_t.
int k1_condense_launch(const BuildParams& P, cudaStream_t st);
int k1_psi_fill_launch(const double* Gs, long long sGs, double* Psi, int nx, int nu, int N, int batch, cudaStream_t st);
int k2k4_assemble_launch(const BuildParams& P, double* schur_ws, long long schur_stride, int sms, size_t smem_optin,
    cudaStream_t st);
// batched FP64 tensor-core GEMM (dgemm_dmma.cu): C = alpha op(A) B + beta C, column-major; returns launches or -cudaError
int dgemm_dmma_launch(int transA, int M, int N, int K, double alpha, const double* A, int lda, long long sA, const double* B, int ldb,
    long long sB, double beta, double* C, int ldc, long long sC, int batch, cudaStream_t st);
// measured DFMA / DMMA ceilings (fp64_peak.cu); 0 or -cudaError
int fp64_peaks_measure(int sms, double* scratch, cudaStream_t st, double* dfma_tflops, double* dmma_tflops);
// K3 alone: materialise the step-size rows of Aeq / Aineq (skipped by builds whose solver evaluates them from the tables)
int k3_fill_rows_launch(const BuildParams& P, cudaStream_t st);
int k4_finalize_launch(const BuildParams& P, int sms, cudaStream_t st);
int k7_results_launch(const BuildParams& P, const double* x, double* control, double* trajectory, cudaStream_t st);

} // namespace cb
