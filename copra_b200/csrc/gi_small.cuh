// gi_small.cuh -- K5+K6 for small QPs (n <= 64 variables): the same dual active-set method as
// gi_solver.cuh, laid out for latency.  One 128-thread CTA (4 warps) per instance, every byte of
// state in shared memory, 4 block barriers per active-set iteration:
//
//   alpha : rank-1 update of J2 left over from the previous ADD  +  slacks of all q constraints and
//           per-warp arg-min of the normalised violation               (independent of J)
//   beta  : pick the most violated constraint; d = J' a   (16 columns per warp, 2 lanes per column)
//   gamma : z = J2 d2 (8 row pairs per warp x 4 column groups, shuffle-reduced), r = S d1, step-length
//           candidates, |d2|^2, z'z, z'a                               (per-warp partials)
//   delta : step lengths, x/u update, reflection vectors for ADD (or the DROP reflection)
//
// Layout rules that make every sweep conflict-free AND 128-bit wide: J and the cached constraint
// rows are column-major with an even leading dimension == 2 (mod 4); a lane owns a PAIR of adjacent
// rows (one LDS.128 / STS.128), a quarter-warp covers 128 contiguous bytes, and a column sweep with
// one lane per column hits 8 distinct 16-byte bank groups per quarter-warp (ld/2 is odd).
#pragma once
#include "common.cuh"
#include "engine.cuh"
#include "gi_solver.cuh"

namespace cb {

constexpr int kSmT = 128; // threads per instance
constexpr int kSmMaxN = 64;
constexpr int kSmMinB = 4; // resident CTAs per SM the register budget is set for (128 registers)

__host__ __device__ inline int ld_vec2(int rows) // rows -> even leading dimension == 2 (mod 4)
{
    const int e = (rows + 1) & ~1;
    return (e % 4 == 2) ? e : e + 2;
}

struct GsLayout {
    int n, n2, ld, lds, mg, mg2, lda;
    size_t oJ, oS, oA, oX, oD, oZ, oW, oV, oR, oU, oBg, oNorm, oLb, oUb, oRow, oRowk, oRed; // doubles
    size_t oIact, oRowmap, oRedI, oActive, oSgn, bytes;                                     // bytes
};

__host__ __device__ inline GsLayout gs_layout(int n, int meq, int m, bool s_smem = true, bool a_smem = true)
{
    GsLayout L;
    L.n = n; L.n2 = (n + 1) & ~1; L.ld = ld_vec2(n); L.lds = odd_ld(n);
    L.mg = meq + m; L.mg2 = (L.mg + 1) & ~1; L.lda = ld_vec2(L.mg > 0 ? L.mg : 2);
    size_t o = 0;
    auto take = [&](size_t cnt) { size_t at = o; o += (cnt + 1) & ~size_t(1); return at; };
    L.oJ = take(size_t(L.ld) * L.n2);
    L.oS = take(s_smem ? size_t(L.lds) * n : 0);
    L.oA = take(L.mg > 0 && a_smem ? size_t(L.lda) * L.n2 : 0);
    L.oX = take(L.n2); L.oD = take(L.n2); L.oZ = take(L.n2); L.oW = take(L.n2); L.oV = take(L.n2);
    L.oR = take(L.n2); L.oU = take(L.n2 + 2); L.oBg = take(L.mg2); L.oNorm = take(L.mg2);
    L.oLb = take(L.n2); L.oUb = take(L.n2); L.oRow = take(L.n2 + 2); L.oRowk = take(L.n2);
    L.oRed = take(64);
    size_t b = o * sizeof(double);
    L.oIact = b; b += sizeof(int) * size_t(L.n2);
    L.oRowmap = b; b += sizeof(int) * size_t(L.n2);
    L.oRedI = b; b += sizeof(int) * 16;
    L.oActive = b; b += size_t(L.mg + 2 * n);
    L.oSgn = b; b += size_t(meq > 0 ? meq : 1);
    L.bytes = (b + 15) & ~size_t(15);
    return L;
}

struct GsWork {
    double *J, *S, *A, *x, *d, *z, *w, *v, *r, *u, *bg, *norm, *lb, *ub, *row, *rowk, *red;
    int *iact, *rowmap, *redi;
    unsigned char* active;
    signed char* sgn;
};

__device__ inline GsWork gs_carve(const GsLayout& L, unsigned char* smem)
{
    GsWork W;
    double* b = reinterpret_cast<double*>(smem);
    W.J = b + L.oJ; W.S = b + L.oS; W.A = b + L.oA; W.x = b + L.oX; W.d = b + L.oD; W.z = b + L.oZ;
    W.w = b + L.oW; W.v = b + L.oV; W.r = b + L.oR; W.u = b + L.oU; W.bg = b + L.oBg; W.norm = b + L.oNorm;
    W.lb = b + L.oLb; W.ub = b + L.oUb; W.row = b + L.oRow; W.rowk = b + L.oRowk; W.red = b + L.oRed;
    W.iact = reinterpret_cast<int*>(smem + L.oIact);
    W.rowmap = reinterpret_cast<int*>(smem + L.oRowmap);
    W.redi = reinterpret_cast<int*>(smem + L.oRedI);
    W.active = smem + L.oActive;
    W.sgn = reinterpret_cast<signed char*>(smem + L.oSgn);
    return W;
}

// Ampere-style asynchronous 8-byte copy global -> shared (no register staging, any number in flight per thread)
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

// out[c] = sum_k M[k + c*ld] * vec[k] over k in [0,n2): warp wq owns columns [16wq,16wq+16), two lanes
// per column (k halves), result valid in lanes < 16.  vec is read with stride `vs` (1 for a vector in
// shared memory, lda for a row of the cached constraint matrix).
template <bool VG> // VG: vec is a row of the un-padded global constraint matrix (n entries; entry n of an odd n does not exist)
__device__ __forceinline__ double gs_col_dot(const double* __restrict__ M, int ld, int n, int n2, const double* __restrict__ vec, int vs)
{
    const int lane = lane_id(), c = (warp_id() << 4) + (lane & 15), h = lane >> 4;
    const int kh = ((n2 >> 1) + 1) & ~1;
    const int k0 = h ? kh : 0, k1 = h ? n2 : kh;
    double s0 = 0.0, s1 = 0.0;
    if (c < n2) {
        const double* col = M + size_t(c) * ld;
#pragma unroll 4
        for (int k = k0; k < k1; k += 2) {
            const double2 j = ld2(col + k);
            s0 += j.x * vec[size_t(k) * vs];
            s1 += j.y * ((VG && k + 1 >= n) ? 0.0 : vec[size_t(k + 1) * vs]);
        }
    }
    s0 += s1;
    s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    return s0;
}

// (z0,z1) for the row pair owned by this lane: sum over columns c in [c0,c1) of M[pair + c*ld]*vec[c].
// warp wq owns row pairs [8wq, 8wq+8) (+32 per pass), 4 column groups per pair; the result is valid in
// the lanes of column group 0 (lane < 8).
__device__ __forceinline__ void gs_rowpair_dot(const double* __restrict__ M, int ld, int pair, bool valid, int c0, int c1,
    const double* __restrict__ vec, double& z0, double& z1)
{
    const int g = lane_id() >> 3;
    z0 = 0.0; z1 = 0.0;
    if (valid) {
        const double* mp = M + 2 * pair;
#pragma unroll 4
        for (int c = c0 + g; c < c1; c += 4) {
            const double2 j = ld2(mp + size_t(c) * ld);
            const double vc = vec[c];
            z0 += j.x * vc;
            z1 += j.y * vc;
        }
    }
    z0 += __shfl_xor_sync(0xffffffffu, z0, 8);
    z1 += __shfl_xor_sync(0xffffffffu, z1, 8);
    z0 += __shfl_xor_sync(0xffffffffu, z0, 16);
    z1 += __shfl_xor_sync(0xffffffffu, z1, 16);
}

// the same product for a row pair of the stacked general rows [Aeq; Aineq] read from global memory (L2): for a
// fixed column the 8 pairs of a warp are 16 consecutive doubles, so every request is one or two 128-byte lines
__device__ __forceinline__ void gs_rowpair_dot_global(const GiView& P, int pair, bool valid, int n, const double* __restrict__ vec,
    double& z0, double& z1)
{
    const int g = lane_id() >> 3;
    z0 = 0.0; z1 = 0.0;
    if (valid) {
        const int i0 = 2 * pair, i1 = i0 + 1, meq = P.meq, m = P.m;
        const double* p0 = (i0 < meq) ? P.Aeq + i0 : P.Aineq + (i0 - meq);
        const int s0 = (i0 < meq) ? meq : m;
        const bool has1 = i1 < meq + m;
        const double* p1 = (i1 < meq) ? P.Aeq + i1 : P.Aineq + (has1 ? i1 - meq : 0);
        const int s1 = (i1 < meq) ? meq : m;
#pragma unroll 4
        for (int c = g; c < n; c += 4) {
            const double a0 = __ldg(p0 + size_t(c) * s0);
            const double a1 = has1 ? __ldg(p1 + size_t(c) * s1) : 0.0;
            const double vc = vec[c];
            z0 += a0 * vc;
            z1 += a1 * vc;
        }
    }
    z0 += __shfl_xor_sync(0xffffffffu, z0, 8);
    z1 += __shfl_xor_sync(0xffffffffu, z1, 8);
    z0 += __shfl_xor_sync(0xffffffffu, z0, 16);
    z1 += __shfl_xor_sync(0xffffffffu, z1, 16);
}

// M[pair + c*ld] -= (w0,w1) * cv[c] for c in [c0,c1): same lane mapping as gs_rowpair_dot
__device__ __forceinline__ void gs_rank1(double* __restrict__ M, int ld, int np, int c0, int c1, const double* __restrict__ wv,
    const double* __restrict__ cv)
{
    const int lane = lane_id(), g = lane >> 3;
    for (int pair = (warp_id() << 3) + (lane & 7); pair < np; pair += 32) {
        const double2 w = ld2(wv + 2 * pair);
        double* mp = M + 2 * pair;
#pragma unroll 4
        for (int c = c0 + g; c < c1; c += 4) {
            double2 j = ld2(mp + size_t(c) * ld);
            const double vc = cv[c];
            j.x -= w.x * vc;
            j.y -= w.y * vc;
            *reinterpret_cast<double2*>(mp + size_t(c) * ld) = j;
        }
    }
}

// Upper Cholesky Q = R'R and J = R^-1, fused: step k of LINPACK dpofa (right-looking) and step k of dpori touch
// disjoint entries (rows > k vs rows <= k of the columns j > k), so both are ONE rank-1 sweep
//     J[i,j] = (i == k ? 0 : J[i,j]) + mult[j] * coef[i],   i <= j,  j > k
// with mult[j] = R[k,j], coef[i] = -J[i,k]/R[k,k] (i < k), 1/R[k,k] (i == k), -R[k,i] (i > k).  Every entry sees
// its updates in LINPACK's order.  128 threads, lane = row pair, warp = column group.
//
// NB pivots per pass (two block barriers per pass): the sweeps of the later pivots of a pass only need the block's rows and
// columns as the earlier pivots leave them, and every thread can form its own entries of those from pre-pass values -- the
// NB x NB diagonal block is factored redundantly in registers, each thread folds its NB coefficients / multipliers, and all
// NB rank-1 sweeps are applied to each entry in one visit (same operands, same order as NB single passes).  That divides the
// barriers and the shared-memory traffic of the factorisation by NB.
template <int NB>
__device__ __forceinline__ bool gs_factor_pass(double* __restrict__ J, int ld, int n, int n2, int k, double* const* coefp, double* const* multp)
{
    const int tid = threadIdx.x, lane = lane_id(), g = tid >> 5;
    // ---- the diagonal block, redundantly in every thread (broadcast loads of pre-pass values) -----------------------------
    double R[NB][NB], inv[NB];
    bool pd = true;
#pragma unroll
    for (int p = 0; p < NB; ++p) {
        double dg = J[(k + p) + size_t(k + p) * ld];
#pragma unroll
        for (int q = 0; q < p; ++q) dg = fma(R[q][p], -R[q][p], dg);
        pd = pd && dg > 0.0;
        inv[p] = rsqrt(dg); // 1 / R[p,p]; rows are scaled by multiplying with it (LINPACK divides by sqrt: same to rounding)
#pragma unroll
        for (int c = p + 1; c < NB; ++c) {
            double v = J[(k + p) + size_t(k + c) * ld];
#pragma unroll
            for (int q = 0; q < p; ++q) v = fma(R[q][c], -R[q][p], v);
            R[p][c] = v * inv[p];
        }
    }
    if (!pd) return false; // uniform: every thread computed the same values
    // ---- this thread's row / column of the block ---------------------------------------------------------------------------
    double cf[NB], mu[NB];
#pragma unroll
    for (int p = 0; p < NB; ++p) { cf[p] = 0.0; mu[p] = 0.0; }
    if (tid < n) {
        if (tid >= k + NB) { // column tid of the block rows: R[k+p, tid]
#pragma unroll
            for (int p = 0; p < NB; ++p) {
                double v = J[(k + p) + size_t(tid) * ld];
#pragma unroll
                for (int q = 0; q < p; ++q) v = fma(mu[q], -R[q][p], v);
                mu[p] = v * inv[p];
                cf[p] = -mu[p];
            }
        } else if (tid < k) { // row tid of the block columns
#pragma unroll
            for (int p = 0; p < NB; ++p) {
                double v = J[tid + size_t(k + p) * ld];
#pragma unroll
                for (int q = 0; q < p; ++q) v = fma(R[q][p], cf[q], v);
                cf[p] = v * (-inv[p]);
            }
        } else { // a block row t = tid - k: below the pivots p < t, the pivot itself at p == t, above the pivots p > t
#pragma unroll
            for (int t = 0; t < NB; ++t) { // compile-time t: everything stays in registers
                if (tid != k + t) continue;
#pragma unroll
                for (int p = 0; p < NB; ++p) {
                    if (p < t) { mu[p] = R[p][t]; cf[p] = -mu[p]; }
                    else if (p == t) cf[p] = inv[p];
                    else {
                        double v = 0.0; // pivot t zeroed this row first
#pragma unroll
                        for (int q = t; q < p; ++q) v = fma(R[q][p], cf[q], v);
                        cf[p] = v * (-inv[p]);
                    }
                }
            }
        }
    }
    if (tid < n2) {
#pragma unroll
        for (int p = 0; p < NB; ++p) { coefp[p][tid] = cf[p]; multp[p][tid] = mu[p]; }
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < NB; ++p)
        if (tid <= k + p) J[tid + size_t(k + p) * ld] = cf[p]; // the finished columns of the inverse
    {
        const int i0 = 2 * lane, i1 = i0 + 1, j0 = k + NB;
        if (i0 < n) {
            double2 c[NB];
#pragma unroll
            for (int p = 0; p < NB; ++p) c[p] = ld2(coefp[p] + i0);
            for (int j = max(j0, i0) + ((g - max(j0, i0)) & 3); j < n; j += 4) {
                double2 a = ld2(J + i0 + size_t(j) * ld);
                const bool second = i1 <= j;
#pragma unroll
                for (int p = 0; p < NB; ++p) {
                    const double m = multp[p][j];
                    a.x = fma(m, c[p].x, (i0 == k + p) ? 0.0 : a.x);
                    if (second) a.y = fma(m, c[p].y, (i1 == k + p) ? 0.0 : a.y);
                }
                *reinterpret_cast<double2*>(J + i0 + size_t(j) * ld) = a;
            }
        }
    }
    __syncthreads();
    return true;
}

// coefp / multp: NB + NB scratch vectors of n2 doubles (16-byte aligned).  NB = 2 and NB = 4 measure the same on C2 / C4
// (1.59 / 5.99 ms: the longer redundant prelude of NB = 4 eats what its fewer barriers save); callers use kSmFacNB.
constexpr int kSmFacNB = 2;
template <int NB>
__device__ inline bool gs_factor_nb(double* __restrict__ J, int ld, int n, int n2, double* const* coefp, double* const* multp)
{
    int k = 0;
    for (; k + NB <= n; k += NB)
        if (!gs_factor_pass<NB>(J, ld, n, n2, k, coefp, multp)) return false;
    if (NB > 2 && k + 2 <= n) {
        if (!gs_factor_pass<2>(J, ld, n, n2, k, coefp, multp)) return false;
        k += 2;
    }
    if (k < n && !gs_factor_pass<1>(J, ld, n, n2, k, coefp, multp)) return false;
    // strict lower triangle := 0 (qpgen2 label 21); pad row/column are zeroed by the caller
    const int lane = lane_id(), g = threadIdx.x >> 5;
    for (int j = g; j < n2; j += kSmT / 32)
        for (int i = j + 1 + lane; i < n2; i += 32) J[i + size_t(j) * ld] = 0.0;
    __syncthreads();
    return true;
}

// All 128 threads call with identical arguments.  Returns the QuadProg fail code.
// AG: the general constraint rows stay in global memory (L2) in the caller's layout instead of a padded copy in
// shared memory; SG: S = R^-1 lives in a per-CTA global workspace `Sg` (odd_ld(n) * n doubles).  Both trade a
// little L2 latency per iteration for 20 KB of shared memory each, i.e. for more resident instances per SM.
template <bool AG, bool SG>
__device__ inline int gs_solve(const GiView& P, const GsLayout& L, GsWork& W, double* __restrict__ Sg, const GiOut& O, double vsmall,
    int max_iter)
{
    const int n = L.n, n2 = L.n2, ld = L.ld, lds = L.lds, meq = P.meq, m = P.m, mg = L.mg, mg2 = L.mg2, lda = L.lda;
    const int q = mg + 2 * n, np = n2 >> 1, npa = mg2 >> 1;
    const int tid = threadIdx.x, lane = lane_id(), wq = warp_id();
    double* __restrict__ J = W.J;
    double* __restrict__ S = SG ? Sg : W.S;
    double* __restrict__ A = W.A;
    // element (i, k) of the stacked general rows [Aeq; Aineq] straight from global memory (AG)
    auto ag = [&](int i, int k) -> double { return (i < meq) ? __ldg(P.Aeq + i + size_t(k) * meq) : __ldg(P.Aineq + (i - meq) + size_t(k) * m); };
    double* redv = W.red;        // [0..4) arg-min values, [4..8) slack of the winner / dd, [8..12) zz, [12..16) za
    int* redi = W.redi;

    // ---- 0. load (zero padded) ---------------------------------------------------------------------
    // Q and the general rows stream in with cp.async (no register staging, every copy of a thread in flight at once); the
    // factorisation starts as soon as Q has landed, the rows are only awaited before their norms are needed
    if (P.Jin) { // receding-horizon re-solve: the factor of this instance is resident, nothing to factor
        for (int idx = tid; idx < ld * n2; idx += kSmT) cp_async8(J + idx, P.Jin + idx);
    } else {
        for (int idx = tid; idx < ld * n2; idx += kSmT) {
            const int i = idx % ld, j = idx / ld;
            if (i < n && j < n) cp_async8(J + idx, P.Q + i + size_t(j) * n);
            else J[idx] = (i == j && i < n2) ? 1.0 : 0.0;
        }
    }
    cp_async_commit();
    if (mg > 0) {
        if (!AG)
            for (int idx = tid; idx < lda * n2; idx += kSmT) {
                const int i = idx % lda, k = idx / lda;
                if (k < n && i < meq) cp_async8(A + idx, P.Aeq + i + size_t(k) * meq);
                else if (k < n && i < mg) cp_async8(A + idx, P.Aineq + (i - meq) + size_t(k) * m);
                else A[idx] = 0.0;
            }
        for (int i = tid; i < mg2; i += kSmT) W.bg[i] = (i < meq) ? P.beq[i] : (i < mg ? P.bineq[i - meq] : 0.0);
    }
    for (int i = tid; i < n2; i += kSmT) {
        const bool in = i < n;
        W.v[i] = in ? -P.c[i] : 0.0;
        W.lb[i] = in ? P.lb[i] : 0.0;
        W.ub[i] = in ? P.ub[i] : 0.0;
        W.u[i] = 0.0; W.x[i] = 0.0; W.d[i] = 0.0; W.z[i] = 0.0; W.w[i] = 0.0; W.r[i] = 0.0;
        W.iact[i] = 0;
        W.rowmap[i] = i;
    }
    if (tid < 2) W.u[n2 + tid] = 0.0;
    for (int i = tid; i < q; i += kSmT) W.active[i] = 0;
    for (int i = tid; i < meq; i += kSmT) W.sgn[i] = 1;
    cp_async_commit();   // group 1: the general rows (possibly empty)
    cp_async_wait<1>();  // group 0 (Q) has landed for this thread ...
    __syncthreads();     // ... and for everyone

    int fail = 0, nact = 0, iter0 = 0, iter1 = 0;
    // the padded diagonal entry (odd n) keeps the factorisation well defined; it is zeroed afterwards
    // (the 8-pivot blocked DMMA factorisation of gi_factor.cuh was measured slower here: C2 2.03 vs 1.78 ms -- at n ~ 50 the
    // per-pivot latency chain dominates, not the sweep)
    if (!P.Jin) { // scratch of the factorisation: vectors that hold nothing yet (red has 64 entries >= n2)
        double* const coefp[4] = { W.x, W.d, W.z, W.w };
        double* const multp[4] = { W.row, W.rowk, W.r, W.red };
        if (!gs_factor_nb<kSmFacNB>(J, ld, n, n2, coefp, multp)) fail = 2;
    }
    cp_async_wait<0>(); // the general rows have landed (also drains the copies before the buffers are reused)
    __syncthreads();
    if (!P.Jin && P.Jflag && tid == 0) *P.Jflag = (fail == 0) ? 1 : 0;
    if (fail == 0) {
        if (n2 > n && !P.Jin) {
            for (int i = tid; i < n2; i += kSmT) { J[n + size_t(i) * ld] = 0.0; J[i + size_t(n) * ld] = 0.0; }
            __syncthreads();
        }
        if (P.Jout) for (int idx = tid; idx < ld * n2; idx += kSmT) P.Jout[idx] = J[idx]; // kept for the next re-solve
        // unconstrained minimiser x = J (J' (-c))
        {
            const double s = gs_col_dot<false>(J, ld, n, n2, W.v, 1);
            if (lane < 16 && (wq << 4) + lane < n2) W.d[(wq << 4) + lane] = s;
        }
        __syncthreads();
        for (int pair = (wq << 3) + (lane & 7); pair < 32; pair += 32) {
            double z0, z1;
            gs_rowpair_dot(J, ld, pair, pair < np, 0, n, W.d, z0, z1);
            if (lane < 8 && pair < np) st2(W.x + 2 * pair, z0, z1);
        }
        // norms of the general rows
        for (int i = tid; i < mg; i += kSmT) {
            double s = 0.0;
            for (int k = 0; k < n; ++k) { const double v = AG ? ag(i, k) : A[i + size_t(k) * lda]; s += v * v; }
            W.norm[i] = 1.0 / sqrt(s); // reciprocal norm: the per-iteration normalisation is a multiply
        }
        __syncthreads();

        bool pending = false; // a rank-1 update of J2 (columns [pc0, n)) waits to be applied in phase alpha
        int pc0 = 0;
        for (;;) {
            ++iter0;
            if (iter0 > max_iter) { fail = 3; break; }
            // ================= alpha: pending rank-1 + all slacks + per-warp arg-min ====================
            if (pending) { gs_rank1(J, ld, np, pc0, n, W.w, W.v); pending = false; }
            MinIdx best; best.v = 0.0; best.i = -1;
            double best_s = 0.0;
            auto consider = [&](int i, double s) {
                if (fabs(s) < vsmall) s = 0.0;
                if (i < meq) {
                    if (s > 0.0) W.sgn[i] = -W.sgn[i];
                    s = -fabs(s);
                }
                if (W.active[i]) s = 0.0;
                if (s < 0.0) {
                    MinIdx c; c.v = (i < mg) ? s * W.norm[i] : s; c.i = i; // bound rows have unit normals
                    const MinIdx nb = better(best, c);
                    if (nb.i != best.i) best_s = s;
                    best = nb;
                }
            };
            for (int base = 0; base < npa; base += 32) { // general rows: products with x, row pairs
                const int pair = base + (wq << 3) + (lane & 7);
                double p0, p1;
                if (AG) gs_rowpair_dot_global(P, pair, pair < npa, n, W.x, p0, p1);
                else gs_rowpair_dot(A, lda, pair, pair < npa, 0, n, W.x, p0, p1);
                if (lane < 8 && pair < npa) {
                    const int i0 = 2 * pair, i1 = i0 + 1;
                    consider(i0, (i0 < meq) ? double(W.sgn[i0]) * (p0 - W.bg[i0]) : W.bg[i0] - p0);
                    if (i1 < mg) consider(i1, (i1 < meq) ? double(W.sgn[i1]) * (p1 - W.bg[i1]) : W.bg[i1] - p1);
                }
            }
            for (int j = tid; j < 2 * n; j += kSmT) { // bound rows: upper (-x_j >= -ub_j) then lower (x_j >= lb_j)
                consider(mg + j, gi_bound_slack(j, n, mg, W.x, W.lb, W.ub, W.active));
            }
            {
                const MinIdx wm = warp_argmin(best);
                const unsigned own = __ballot_sync(0xffffffffu, wm.i >= 0 && best.i == wm.i);
                const double ws = __shfl_sync(0xffffffffu, best_s, own ? (__ffs(own) - 1) : 0);
                if (lane == 0) { redv[wq] = wm.v; redi[wq] = wm.i; redv[4 + wq] = ws; }
            }
            __syncthreads();
            // ================= beta: select; d = J' a ====================================================
            MinIdx sel; sel.v = redv[0]; sel.i = redi[0];
            double s_nvl = redv[4];
#pragma unroll
            for (int k = 1; k < 4; ++k) {
                MinIdx c; c.v = redv[k]; c.i = redi[k];
                const MinIdx nb = better(sel, c);
                if (nb.i != sel.i) s_nvl = redv[4 + k];
                sel = nb;
            }
            if (sel.i < 0) break; // optimal
            const int nvl = sel.i;
            int bj = -1;
            double asign = -1.0; // general row: a = asign * A[nvl,:] ; bound row: a = asign * e_bj
            if (nvl < meq) asign = double(W.sgn[nvl]);
            else if (nvl >= mg) { bj = nvl - mg; if (bj >= n) { bj -= n; asign = 1.0; } }
            // row nvl of the general constraints: (pointer, element stride)
            const double* arow = A + nvl;
            int astr = lda;
            if (AG && bj < 0) {
                if (nvl < meq) { arow = P.Aeq + nvl; astr = meq; } else { arow = P.Aineq + (nvl - meq); astr = m; }
            }

            for (;;) { // label 55
                if (bj >= 0) {
                    if (tid < n2) W.d[tid] = (tid < n) ? asign * J[bj + size_t(tid) * ld] : 0.0;
                } else {
                    const double s = gs_col_dot<AG>(J, ld, n, n2, arow, astr);
                    if (lane < 16 && (wq << 4) + lane < n2) W.d[(wq << 4) + lane] = asign * s;
                }
                __syncthreads();
                // ============= gamma: z = J2 d2, r = S d1, candidates and norms ==========================
                double zz = 0.0, za = 0.0;
                {
                    const int pair = (wq << 3) + (lane & 7);
                    double z0, z1;
                    gs_rowpair_dot(J, ld, pair, pair < np, nact, n, W.d, z0, z1);
                    if (lane < 8 && pair < np) {
                        st2(W.z + 2 * pair, z0, z1);
                        zz = z0 * z0 + z1 * z1;
                        const int i0 = 2 * pair;
                        if (bj >= 0) za = (i0 == bj) ? asign * z0 : ((i0 + 1 == bj) ? asign * z1 : 0.0);
                        else za = asign * (z0 * arow[size_t(i0) * astr] + ((AG && i0 + 1 >= n) ? 0.0 : z1 * arow[size_t(i0 + 1) * astr]));
                    }
                }
                MinIdx tc; tc.v = 0.0; tc.i = -1;
                {
                    // r = S d1: two lanes per active row (even / odd k), i.e. all four warps share the sweep -- with S in
                    // L2 (SG) this is the longest chain of the phase.  s0 (even k) + s1 (odd k) as before.
                    const int row = tid >> 1, half = tid & 1;
                    double sh = 0.0;
                    if (row < nact) {
                        const double* srow = S + W.rowmap[row];
                        for (int k = half; k < nact; k += 2) sh += srow[size_t(k) * lds] * W.d[k];
                    }
                    const double other = __shfl_xor_sync(0xffffffffu, sh, 1);
                    if (half == 0 && row < nact) {
                        const double s0 = sh + other;
                        W.r[row] = s0;
                        if (W.iact[row] - 1 >= meq && s0 > 0.0) { tc.v = W.u[row] / s0; tc.i = row; }
                    }
                }
                double dd = 0.0;
                if (tid >= nact && tid < n) { const double dj = W.d[tid]; dd = dj * dj; }
                tc = warp_argmin(tc);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) dd += __shfl_xor_sync(0xffffffffu, dd, o);
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) { // zz / za live in lanes 0..7 only (zeros elsewhere): three steps reach lane 0
                    zz += __shfl_xor_sync(0xffffffffu, zz, o);
                    za += __shfl_xor_sync(0xffffffffu, za, o);
                }
                if (lane == 0) { redv[wq] = tc.v; redi[wq] = tc.i; redv[4 + wq] = dd; redv[8 + wq] = zz; redv[12 + wq] = za; }
                __syncthreads();
                // ============= delta: step lengths, x / u, reflection vectors ============================
                MinIdx t1m; t1m.v = redv[0]; t1m.i = redi[0];
#pragma unroll
                for (int k = 1; k < 4; ++k) { MinIdx c; c.v = redv[k]; c.i = redi[k]; t1m = better(t1m, c); }
                dd = (redv[4] + redv[5]) + (redv[6] + redv[7]);
                zz = (redv[8] + redv[9]) + (redv[10] + redv[11]);
                za = (redv[12] + redv[13]) + (redv[14] + redv[15]);
                const bool t1inf = t1m.i < 0;
                const double t1 = t1m.v;
                const int it1 = t1m.i;

                bool do_drop = false;
                if (fabs(zz) <= vsmall) {
                    if (t1inf) { fail = 1; break; }
                    if (tid < nact) W.u[tid] -= t1 * W.r[tid];
                    if (tid == 0) W.u[nact] += t1;
                    do_drop = true;
                } else {
                    double tt = -s_nvl / za;
                    bool t2min = true;
                    if (!t1inf && t1 < tt) { tt = t1; t2min = false; }
                    if (tid < n) W.x[tid] += tt * W.z[tid];
                    if (tid < nact) W.u[tid] -= tt * W.r[tid];
                    if (tid == 0) W.u[nact] += tt;
                    if (t2min) {
                        // ---- ADD: H d2 = delta e1 ; w = tau (z - delta J[:,nact]) ; v = d2 - delta e1 -----
                        const double d0 = W.d[nact];
                        // sigma = |d2| and 1/sigma from one reciprocal square root (the reflection only needs them to rounding:
                        // any tau, delta pair consistent to a few ulp keeps J orthogonal to rounding like the update itself)
                        const double rsig = rsqrt(dd);
                        const double sigma = dd * rsig;
                        const double delta = (d0 >= 0.0) ? -sigma : sigma;
                        const double tau = 1.0 / (sigma * (sigma + fabs(d0)));
                        if (tid < n2) {
                            W.w[tid] = (tid < n) ? tau * (W.z[tid] - delta * J[tid + size_t(nact) * ld]) : 0.0;
                            W.v[tid] = (tid == nact) ? d0 - delta : W.d[tid];
                        }
                        const int newrow = W.rowmap[nact];
                        const double invd = (d0 >= 0.0) ? -rsig : rsig; // 1 / delta
                        if (tid < nact) {
                            S[W.rowmap[tid] + size_t(nact) * lds] = -W.r[tid] * invd;
                            S[newrow + size_t(tid) * lds] = 0.0;
                        }
                        if (tid == 0) {
                            S[newrow + size_t(nact) * lds] = invd;
                            W.iact[nact] = nvl + 1;
                            W.active[nvl] = 1;
                        }
                        pending = true;
                        pc0 = nact;
                        ++nact;
                        __syncthreads();
                        break; // -> alpha
                    } else {
                        // partial step: refresh s_nvl at the new x (equality sign rule included)
                        __syncthreads(); // x complete
                        double s;
                        if (bj >= 0) s = (asign < 0.0) ? W.ub[bj] - W.x[bj] : W.x[bj] - W.lb[bj];
                        else {
                            double acc = (tid < n) ? arow[size_t(tid) * astr] * W.x[tid] : 0.0;
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                            if (lane == 0) redv[wq] = acc;
                            __syncthreads();
                            acc = (redv[0] + redv[1]) + (redv[2] + redv[3]);
                            s = (nvl < meq) ? asign * (acc - W.bg[nvl]) : W.bg[nvl] - acc;
                        }
                        if (nvl < meq) {
                            if (s > 0.0) { asign = -asign; if (tid == 0) W.sgn[nvl] = -W.sgn[nvl]; }
                            s = -fabs(s);
                        }
                        s_nvl = s;
                        do_drop = true;
                    }
                }
                if (do_drop) {
                    // ---- DROP the it1-th active constraint --------------------------------------------------
                    __syncthreads(); // u updates visible; redv free
                    const int p = it1;
                    const int dropped = (tid == 0) ? W.iact[p] - 1 : 0; // used by thread 0 only, which also clears iact below
                    const int prow = W.rowmap[p];
                    if (nact > 1) {
                        // v = row p of S ; rho = |v| ; w = v - gamma e_last ; tw = tau w
                        double vk = 0.0;
                        if (tid < nact) vk = S[prow + size_t(tid) * lds];
                        double vv = vk * vk;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) vv += __shfl_xor_sync(0xffffffffu, vv, o);
                        if (lane == 0) redv[wq] = vv;
                        if (tid == nact - 1) redv[4] = vk;
                        __syncthreads();
                        const double rho = sqrt((redv[0] + redv[1]) + (redv[2] + redv[3]));
                        const double vl = redv[4];
                        const double gamma = (vl >= 0.0) ? -rho : rho;
                        const double tau = 1.0 / (rho * (rho + fabs(vl)));
                        if (tid < n2) {
                            const double wk = (tid < nact) ? ((tid == nact - 1) ? vk - gamma : vk) : 0.0;
                            W.v[tid] = wk;       // w
                            W.d[tid] = tau * wk; // tau w   (d is recomputed at label 55)
                        }
                        __syncthreads();
                        // J1 w (row pairs of J) and S w (active rows)
                        {
                            const int pair = (wq << 3) + (lane & 7);
                            double z0, z1;
                            gs_rowpair_dot(J, ld, pair, pair < np, 0, nact, W.v, z0, z1);
                            if (lane < 8 && pair < np) st2(W.w + 2 * pair, z0, z1);
                        }
                        double sw = 0.0;
                        if (tid < nact) {
                            const double* srow = S + W.rowmap[tid];
                            for (int k = 0; k < nact; ++k) sw += srow[size_t(k) * lds] * W.v[k];
                        }
                        __syncthreads();
                        gs_rank1(J, ld, np, 0, nact, W.w, W.d);
                        if (tid < nact && tid != p) {
                            double* srow = S + W.rowmap[tid];
                            for (int k = 0; k < nact - 1; ++k) srow[size_t(k) * lds] -= sw * W.d[k];
                        }
                        // close the gap at position p in u / iact / rowmap
                        double uu = 0.0; int ia = 0, rm = 0;
                        const bool mv = tid >= p && tid < nact - 1;
                        if (mv) { uu = W.u[tid + 1]; ia = W.iact[tid + 1]; rm = W.rowmap[tid + 1]; }
                        __syncthreads();
                        if (mv) { W.u[tid] = uu; W.iact[tid] = ia; W.rowmap[tid] = rm; }
                        if (tid == 0) W.rowmap[nact - 1] = prow;
                        __syncthreads();
                    }
                    if (tid == 0) {
                        W.u[nact - 1] = W.u[nact];
                        W.u[nact] = 0.0;
                        W.iact[nact - 1] = 0;
                        W.active[dropped] = 0;
                    }
                    --nact;
                    ++iter1;
                    __syncthreads();
                    continue; // label 55
                }
            }
            if (fail != 0) break;
        }
    }
    __syncthreads();
    if (O.x) for (int i = tid; i < n; i += kSmT) O.x[i] = (fail == 2) ? 0.0 : W.x[i];
    if (O.iact) for (int i = tid; i < n; i += kSmT) O.iact[i] = (i < nact) ? W.iact[i] : 0;
    if (tid == 0) {
        if (O.status) *O.status = fail;
        if (O.iters) { O.iters[0] = iter0; O.iters[1] = iter1; }
        if (O.nact) *O.nact = nact;
    }
    return fail;
}

} // namespace cb
