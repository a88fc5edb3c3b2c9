// dgemm_dmma.cu -- batched FP64 GEMM on the tensor cores: the dense contraction of QP assembly for
// FULL-SIZE (autoSpan'd) entries, where the cost / constraint matrices are dense R x X blocks:
//   T = M Psi (+ N)                        reference src/costFunctions.cpp:65,197   (NN)
//   Q += T' W T, E = (M Phi)' W T          src/costFunctions.cpp:66-69,199-202      (TN)
//   A = E Psi (+ G), Y = E Phi             src/constraints.cpp:68-72,199-203        (NN)
// sm_100a has no f64 kind in tcgen05; FP64 MMA is the warp-level mma.sync m8n8k4 (SASS DMMA.8x8x4).
//
//   C[b] = alpha * op(A[b]) * B[b] + beta * C[b],   op(A) = A (M x K) or A' (A stored K x M); column-major.
//
// CTA tile 64 x 64, K tile 16, 8 warps as 4 (M) x 2 (N), each warp 16 x 32 = 2 x 4 DMMA tiles.
// Shared-memory layouts are chosen so that every fragment load is conflict-free for 64-bit words (a
// half-warp must hit 16 distinct 8-byte banks): A as [k][m] with ld 68 (op N) or [m][k] with ld 20
// (op T), B as [n][k] with ld 20 -- 68 = 4 (mod 16) and 20 = 4 (mod 16).
// The A tile of the NN case is a set of contiguous 512-byte column segments: when the operand is
// 16-byte aligned they are fetched by the TMA engine (cp.async.bulk.shared.global + mbarrier
// complete_tx) and overlap with the B tile's register-staged loads.
#include "launch.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace cb {

namespace {

constexpr int TM = 64, TN = 64, TK = 16;
constexpr int LDA_N = 68; // A stored [k][m]
constexpr int LDK = 20;   // [m][k] / [n][k]

__device__ __forceinline__ void dmma8x8x4(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct GemmArgs {
    int transA, M, N, K, batch;
    double alpha, beta;
    const double* A; int lda; long long sA;
    const double* B; int ldb; long long sB;
    double* C; int ldc; long long sC;
};

template <bool TRANSA> __global__ void __launch_bounds__(256) dgemm_dmma_kernel(const GemmArgs g)
{
    __shared__ __align__(16) double As[TRANSA ? TM * LDK : TK * LDA_N];
    __shared__ __align__(16) double Bs[TN * LDK];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1; // 4 x 2 warps
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    const long long b = blockIdx.z;
    const double* A = g.A + b * g.sA;
    const double* B = g.B + b * g.sB;
    double* C = g.C + b * g.sC;
    const int M = g.M, N = g.N, K = g.K;

    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // TMA is usable for the NN A tile when every 64-double column segment is in range and 16-byte aligned
    const bool tma_ok = !TRANSA && (m0 + TM <= M) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((g.lda & 1) == 0);
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();
    uint32_t phase = 0;

    for (int k0 = 0; k0 < K; k0 += TK) {
        // ---- A tile ----
        if (TRANSA) {
            // As[m][k] = A[(k0+k) + (m0+m) * lda] : k contiguous in global
            for (int idx = tid; idx < TM * TK; idx += 256) {
                const int k = idx % TK, m = idx / TK;
                const int gm = m0 + m, gk = k0 + k;
                As[m * LDK + k] = (gm < M && gk < K) ? A[gk + (long long)gm * g.lda] : 0.0;
            }
        } else if (tma_ok) {
            if (tid == 0) {
                const int kk = min(TK, K - k0);
                mbar_expect_tx(&bar, uint32_t(kk) * TM * sizeof(double));
                for (int k = 0; k < kk; ++k) tma_load_1d(&As[k * LDA_N], A + m0 + (long long)(k0 + k) * g.lda, TM * sizeof(double), &bar);
            }
            if (k0 + TK > K)
                for (int idx = tid; idx < TM * TK; idx += 256) {
                    const int m = idx % TM, k = idx / TM;
                    if (k0 + k >= K) As[k * LDA_N + m] = 0.0;
                }
        } else {
            // As[k][m] = A[(m0+m) + (k0+k) * lda] : m contiguous in global
            for (int idx = tid; idx < TM * TK; idx += 256) {
                const int m = idx % TM, k = idx / TM;
                const int gm = m0 + m, gk = k0 + k;
                As[k * LDA_N + m] = (gm < M && gk < K) ? A[gm + (long long)gk * g.lda] : 0.0;
            }
        }
        // ---- B tile: Bs[n][k] = B[(k0+k) + (n0+n) * ldb] ----
        for (int idx = tid; idx < TN * TK; idx += 256) {
            const int k = idx % TK, n = idx / TK;
            const int gn = n0 + n, gk = k0 + k;
            Bs[n * LDK + k] = (gn < N && gk < K) ? B[gk + (long long)gn * g.ldb] : 0.0;
        }
        if (!TRANSA && tma_ok) { mbar_wait(&bar, phase); phase ^= 1; }
        __syncthreads();
        // ---- 4 k-steps of DMMA ----
#pragma unroll
        for (int ks = 0; ks < TK; ks += 4) {
            double af[2], bf[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int m = wm * 16 + i * 8 + (lane >> 2), k = ks + (lane & 3);
                af[i] = TRANSA ? As[m * LDK + k] : As[k * LDA_N + m];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = wn * 32 + j * 8 + (lane >> 2), k = ks + (lane & 3);
                bf[j] = Bs[n * LDK + k];
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();
    }
    // ---- epilogue ----
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int m = m0 + wm * 16 + i * 8 + (lane >> 2);
                const int n = n0 + wn * 32 + j * 8 + 2 * (lane & 3) + e;
                if (m < M && n < N) {
                    double* cp = C + m + (long long)n * g.ldc;
                    const double v = g.alpha * acc[i][j][e];
                    *cp = (g.beta == 0.0) ? v : v + g.beta * (*cp);
                }
            }
}

} // namespace

int dgemm_dmma_launch(int transA, int M, int N, int K, double alpha, const double* A, int lda, long long sA, const double* B, int ldb,
    long long sB, double beta, double* C, int ldc, long long sC, int batch, cudaStream_t st)
{
    if (M <= 0 || N <= 0 || batch <= 0) return 0;
    int launches = 0;
    for (int b0 = 0; b0 < batch; b0 += 65535) {
        const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
        GemmArgs g;
        g.transA = transA; g.M = M; g.N = N; g.K = K; g.batch = nb; g.alpha = alpha; g.beta = beta;
        g.A = A + (long long)b0 * sA; g.lda = lda; g.sA = sA;
        g.B = B + (long long)b0 * sB; g.ldb = ldb; g.sB = sB;
        g.C = C + (long long)b0 * sC; g.ldc = ldc; g.sC = sC;
        dim3 grid((M + TM - 1) / TM, (N + TN - 1) / TN, nb);
        if (transA) dgemm_dmma_kernel<true><<<grid, 256, 0, st>>>(g);
        else dgemm_dmma_kernel<false><<<grid, 256, 0, st>>>(g);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return -int(e);
        ++launches;
    }
    return launches;
}

} // namespace cb
