// gi_factor.cuh -- blocked Cholesky + inverse for ONE CTA:  J (n x n, column-major, shared OR global memory)
// holds Q on entry and J = R^-1 (Q = R'R, strict lower triangle zero) on exit -- what qpgen2 computes with
// LINPACK dpofa + dpori before its first iteration.
//
// Per pivot p the two LINPACK loops are one fused sweep
//     J[i,j] = (i == p ? 0 : J[i,j]) + R[p,j] * coef_p[i]        (j > p, i <= j)
//     coef_p[i] = -J[i,p] / R[p,p] (i < p),  1 / R[p,p] (i == p),  -R[p,i] (i > p)
// Pivots are taken 8 at a time: the 8 rows of the block are copied to a shared-memory panel and factored there
// (the only part with a barrier per pivot), every row folds its 8 coefficients from the finished panel, and the
// trailing update of ALL rows is a (n x 8) x (8 x cols) product on the FP64 tensor cores (DMMA.8x8x4) that
// streams J exactly once per block -- with J in global memory (n > ~160) that is n^3/8 + O(n^2) words of traffic
// instead of the ~2n^3/3 of the pivot-by-pivot sweeps, and 2 + 16 block barriers per 8 pivots instead of 32.
// Entries below the diagonal are updated too (never read, zeroed at the end).  Each entry sees its updates in pivot
// order; the DMMA accumulates the 4 products of a k-step internally, so J agrees with the rank-1 formulation to
// rounding, not bit for bit.
#pragma once
#include "common.cuh"

namespace cb {

constexpr int kFacNB = 8;

// shared-memory scratch the factorisation needs, in doubles
__host__ __device__ inline size_t gi_factor_scratch(int n) { return size_t(kFacNB) * (n + 2) + size_t(n) * kFacNB + kFacNB; }

__device__ inline bool gi_factor_blocked(double* __restrict__ J, int ld, int n, double* __restrict__ scratch)
{
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
    const int pstride = n + 2;
    double* panel = scratch;                               // 8 x (n+2): [pp][j] = R[p,j] (j >= p), [pp][n] = pivot ok
    double* coefb = panel + size_t(kFacNB) * pstride;      // n x 8    : coef_p[i]
    double* invb = coefb + size_t(n) * kFacNB;             // 8        : 1 / R[p,p]
    bool pd = true;
    for (int k0 = 0; k0 < n; k0 += kFacNB) {
        const int k1 = min(k0 + kFacNB, n), nb = k1 - k0;
        // ---- panel: the block rows, columns >= k0 (8 consecutive rows of a column are 64 contiguous bytes) --------
        for (int j = k0 + tid; j < n; j += T)
            for (int pp = 0; pp < nb; ++pp) panel[size_t(pp) * pstride + j] = J[(k0 + pp) + size_t(j) * ld];
        __syncthreads();
        bool ok = true;
        for (int p = k0; p < k1; ++p) {
            const int pp = p - k0;
            double* prow = panel + size_t(pp) * pstride;
            const double akk = prow[p];
            ok = ok && akk > 0.0;
            const double rkk = ok ? sqrt(akk) : 1.0;
            __syncthreads(); // everyone has read the pivot before it is overwritten
            for (int j = p + tid; j < n; j += T) prow[j] = (j == p) ? rkk : prow[j] / rkk;
            if (tid == 0) prow[n] = ok ? 1.0 : 0.0;
            __syncthreads();
            // the rest of the block rows: A[i,j] -= R[p,i] R[p,j], p < i < k1, j >= i (one thread per column)
            const int rows = k1 - 1 - p;
            if (rows > 0) {
                for (int c_ = p + 1 + tid; c_ < n; c_ += T) {
                    const double mc = prow[c_];
                    const int rmax = min(rows, c_ - p);
                    for (int r_ = 0; r_ < rmax; ++r_) {
                        double* e = panel + size_t(pp + 1 + r_) * pstride + c_;
                        *e = fma(mc, -prow[p + 1 + r_], *e);
                    }
                }
                __syncthreads();
            }
        }
        for (int pp = 0; pp < nb; ++pp) pd = pd && panel[size_t(pp) * pstride + n] != 0.0;
        if (!pd) break; // uniform: every thread reads the same flags
        if (tid < nb) invb[tid] = 1.0 / panel[size_t(tid) * pstride + k0 + tid];
        __syncthreads();
        // ---- coefficients of every row, pivot by pivot, and the finished columns k0..k1-1 ---------------------------
        for (int i = tid; i < n; i += T) {
            double cf[kFacNB];
            const int q0 = (i >= k0) ? i - k0 : 0; // a block row: its own pivot zeroes it, earlier pivots are in the panel
#pragma unroll
            for (int pp = 0; pp < kFacNB; ++pp) {
                if (pp < nb) {
                    const int p = k0 + pp;
                    double c;
                    if (i > p) c = -panel[size_t(pp) * pstride + i];
                    else if (i == p) c = invb[pp];
                    else {
                        double a = J[i + size_t(p) * ld];
#pragma unroll
                        for (int qq = 0; qq < kFacNB; ++qq)
                            if (qq >= q0 && qq < pp) a = fma(panel[size_t(qq) * pstride + p], cf[qq], (i == k0 + qq) ? 0.0 : a);
                        c = a * (-invb[pp]);
                    }
                    cf[pp] = c;
                    coefb[i + size_t(pp) * n] = c;
                    if (i <= p) J[i + size_t(p) * ld] = c;
                } else cf[pp] = 0.0;
            }
        }
        __syncthreads();
        // ---- trailing update of the columns >= k1: DMMA for every 8-row tile except the block's own -----------------
        {
            const int MT = (n + 7) >> 3, NT = (n - k1 + 7) >> 3, NT4 = (NT + 3) >> 2, skip = k0 >> 3;
            for (int t = warp; t < MT * NT4; t += nwarp) { // 4 column tiles per step: 8 independent loads in flight per lane
                const int mt = t % MT, g4 = t / MT;
                if (mt == skip) continue;
                const int ar = (mt << 3) + (lane >> 2), ak = lane & 3;
                const double a0 = (ar < n && ak < nb) ? coefb[ar + size_t(ak) * n] : 0.0;
                const double a1 = (ar < n && ak + 4 < nb) ? coefb[ar + size_t(ak + 4) * n] : 0.0;
                double c0_[4], c1_[4], b0[4], b1[4];
#pragma unroll
                for (int u_ = 0; u_ < 4; ++u_) {
                    const int cb_ = k1 + (((g4 << 2) + u_) << 3), cc = cb_ + 2 * (lane & 3);
                    c0_[u_] = (ar < n && cc < n) ? J[ar + size_t(cc) * ld] : 0.0;
                    c1_[u_] = (ar < n && cc + 1 < n) ? J[ar + size_t(cc + 1) * ld] : 0.0;
                    const int bc = min(cb_ + (lane >> 2), n - 1);
                    b0[u_] = panel[size_t(ak) * pstride + bc];
                    b1[u_] = panel[size_t(ak + 4) * pstride + bc];
                }
#pragma unroll
                for (int u_ = 0; u_ < 4; ++u_) {
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0_[u_]), "+d"(c1_[u_]) : "d"(a0), "d"(b0[u_]));
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0_[u_]), "+d"(c1_[u_]) : "d"(a1), "d"(b1[u_]));
                }
#pragma unroll
                for (int u_ = 0; u_ < 4; ++u_) {
                    const int cc = k1 + (((g4 << 2) + u_) << 3) + 2 * (lane & 3);
                    if (ar < n && cc < n) J[ar + size_t(cc) * ld] = c0_[u_];
                    if (ar < n && cc + 1 < n) J[ar + size_t(cc + 1) * ld] = c1_[u_];
                }
            }
            // the block rows: pivot k0+rr zeroes row rr first, later pivots add their term (one thread per column)
            for (int c_ = k1 + tid; c_ < n; c_ += T) {
                for (int rr = 0; rr < nb; ++rr) {
                    double a = 0.0;
                    for (int pp = rr; pp < nb; ++pp) a = fma(panel[size_t(pp) * pstride + c_], coefb[(k0 + rr) + size_t(pp) * n], a);
                    J[(k0 + rr) + size_t(c_) * ld] = a;
                }
            }
        }
        __syncthreads();
    }
    if (!pd) return false;
    for (int idx = tid; idx < n * n; idx += T) { // strict lower triangle := 0 (qpgen2 label 21)
        const int i = idx % n, j = idx / n;
        if (i > j) J[i + size_t(j) * ld] = 0.0;
    }
    __syncthreads();
    return true;
}

} // namespace cb
