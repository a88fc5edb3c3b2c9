// gi_solver.cuh -- K5+K6: batched Goldfarb-Idnani dual active-set QP solver, one CTA per instance.
//
// Replaces QuadProgDenseSolver::SI_solve (reference src/QuadProgSolver.cpp:54-72) and the external
// Eigen::QuadProgDense::solve -> qpgen2 behind it (SURVEY.md 3.3 is the specification):
//     min 1/2 x'Qx + c'x   s.t.  Aeq x = beq, Aineq x <= bineq, lb <= x <= ub
// mapped onto quadprog's  min 1/2 x'Dx - d'x, a_i'x >= b_i  with the constraint index space
//     [0,meq) equalities | [meq,meq+m) -Aineq rows | [.., +n) upper bounds (-e_j) | [.., +n) lower (+e_j)
// The 2n bound rows are never materialised (K5) but keep their indices so selection order and the
// reported active set match QuadProg's.
//
// B200-first formulation.  The iterates (x, u, active set, step lengths) are those of qpgen2 in exact
// arithmetic; what changes is the linear algebra that carries the factorisation, chosen so that every
// step is a GEMV, a rank-1 update or a reduction over a whole CTA (no serial chains):
//   * J = L^-T Q (n x n) lives in shared memory (column-major, odd leading dimension => conflict-free
//     sweeps along rows and along columns) or, when it does not fit one SM, in a global workspace.
//   * qpgen2 keeps the triangular factor R of the active normals and back-substitutes r = R^-1 d1 (a
//     length-nact serial chain).  Here S = R^-1 is kept explicitly: r = S d1 is a parallel mat-vec.
//     S need not stay triangular, so orthogonal updates are free to be Householder reflections:
//   * ADD  (qpgen2 labels 150-180: a chain of n-nact Givens rotations folding d2 into its first
//     component): one reflection H d2 = delta e1, J2 <- J2 - tau (J2 v) v' with J2 v = z - delta J[:,nact]
//     (z is already known), a perfectly parallel rank-1 update; S gains the column [-r/delta; 1/delta].
//   * DROP (labels 700-799: Givens sweep over the Hessenberg factor): one reflection whose last column
//     is the dropped row of S; J1 and S get rank-1 updates, the last column joins J2.
#pragma once
#include "common.cuh"
#include "gi_factor.cuh"
#ifdef GC_PROFILE
#include <cstdio>
#endif
#include "engine.cuh"

namespace cb {

struct GiView { // one instance, device pointers, dense column-major
    int n, meq, m;
    const double* Q;     // n x n (upper triangle used, like dpofa)
    const double* c;     // n
    const double* Aeq;   // meq x n, ld meq
    const double* beq;   // meq
    const double* Aineq; // m x n, ld m
    const double* bineq; // m
    const double* lb;    // n
    const double* ub;    // n
    const double* Jin = nullptr;  // gi_small: a cached factor to load instead of factoring Q (see GiBatch::jcache)
    double* Jout = nullptr;       // gi_small: where to store the factor
    int* Jflag = nullptr;
};

struct GiOut {
    double* x;   // n
    int* status; // 1
    int* iters;  // 2
    int* nact;   // 1
    int* iact;   // n (1-based, 0 padded)
};

// Shared-memory carve-up; identical on host (to size the launch) and device.
struct GiLayout {
    int n, meq, m, ldj, lds, lda, threads;
    int j_smem, s_smem, a_smem;
    size_t oJ, oS, oA, oX, oD, oZ, oAv, oR, oU, oW, oV, oRow, oNorm, oLb, oUb, oSl, oPart, oRed; // doubles
    size_t oIact, oRowmap, oRedI, oActive, oSgn, bytes;                                          // bytes
};

__host__ __device__ inline int round32(int v) { return (v + 31) & ~31; }

__host__ __device__ inline GiLayout gi_layout(int n, int meq, int m, int threads, int j_smem, int s_smem, int a_smem)
{
    GiLayout L;
    L.n = n; L.meq = meq; L.m = m; L.threads = threads;
    L.ldj = odd_ld(n); L.lds = odd_ld(n); L.lda = odd_ld(meq + m);
    L.j_smem = j_smem; L.s_smem = s_smem; L.a_smem = a_smem;
    size_t o = 0;
    L.oJ = o; if (j_smem) o += size_t(L.ldj) * n;
    L.oS = o; if (s_smem) o += size_t(L.lds) * n;
    L.oA = o; if (a_smem) o += size_t(L.lda) * n;
    L.oAv = o; o += n;
    L.oU = o; o += n + 1;
    L.oNorm = o; o += meq + m;
    L.oLb = o; o += n;
    L.oUb = o; o += n;
    L.oRed = o; o += 4 * kMaxWarps;
    // the vectors below hold nothing while the factorisation runs: its panel / coefficient scratch aliases them (and
    // extends past them when n is large)
    L.oX = o; o += n;
    L.oD = o; o += n;
    L.oZ = o; o += n;
    L.oR = o; o += n;
    L.oW = o; o += n;
    L.oV = o; o += n;
    L.oRow = o; o += n + 1;
    L.oSl = o; o += meq + m;                    // products of the general rows with x
    L.oPart = o; o += 2 * size_t(threads) + 64; // partial sums of split dot products
    if (o < L.oX + gi_factor_scratch(n)) o = L.oX + gi_factor_scratch(n);
    size_t b = o * sizeof(double);
    L.oIact = b; b += sizeof(int) * size_t(n);
    L.oRowmap = b; b += sizeof(int) * size_t(n);
    L.oRedI = b; b += sizeof(int) * kMaxWarps;
    L.oActive = b; b += size_t(meq + m + 2 * n);
    L.oSgn = b; b += size_t(meq > 0 ? meq : 1);
    L.bytes = (b + 15) & ~size_t(15);
    return L;
}

struct GiWork { // resolved pointers for one CTA
    double *J, *S, *A; // A: cached [Aeq; Aineq] rows, (meq+m) x n col-major ld lda (or nullptr)
    double *x, *d, *z, *av, *r, *u, *w, *v, *row, *norm, *lb, *ub, *sl, *part, *red;
    int *iact, *rowmap, *redi;
    unsigned char* active;
    signed char* sgn;
    int ldj, lds, lda;
};

__device__ inline GiWork gi_carve(const GiLayout& L, unsigned char* smem, double* gJ, double* gS)
{
    GiWork W;
    double* base = reinterpret_cast<double*>(smem);
    W.J = L.j_smem ? base + L.oJ : gJ;
    W.S = L.s_smem ? base + L.oS : gS;
    W.A = L.a_smem ? base + L.oA : nullptr;
    W.x = base + L.oX; W.d = base + L.oD; W.z = base + L.oZ; W.av = base + L.oAv; W.r = base + L.oR;
    W.u = base + L.oU; W.w = base + L.oW; W.v = base + L.oV; W.row = base + L.oRow;
    W.norm = base + L.oNorm; W.lb = base + L.oLb; W.ub = base + L.oUb; W.sl = base + L.oSl;
    W.part = base + L.oPart; W.red = base + L.oRed;
    W.iact = reinterpret_cast<int*>(smem + L.oIact);
    W.rowmap = reinterpret_cast<int*>(smem + L.oRowmap);
    W.redi = reinterpret_cast<int*>(smem + L.oRedI);
    W.active = smem + L.oActive;
    W.sgn = reinterpret_cast<signed char*>(smem + L.oSgn);
    W.ldj = L.ldj; W.lds = L.lds; W.lda = L.lda;
    return W;
}

// ---- CTA-wide building blocks (all threads call with identical arguments) -------------------------
// Visit every (r, c) of a rows x [c0,c1) tile once: lanes run along r (the unit-stride dimension of a
// column-major matrix), the column range is dealt out to G = T / round32(rows) thread groups.
template <class F> __device__ __forceinline__ void tile_rc(int rows, int c0, int c1, F f)
{
    const int tid = threadIdx.x, T = blockDim.x;
    const int rp = max(32, round32(rows));
    if (T >= rp) {
        const int G = T / rp, g = tid / rp, r = tid - g * rp;
        if (g < G && r < rows)
            for (int c = c0 + g; c < c1; c += G) f(r, c);
    } else {
        for (int r = tid; r < rows; r += T)
            for (int c = c0; c < c1; ++c) f(r, c);
    }
}

// "row dots": out[r] = sum_{c in [c0,c1)} M[r + c*ld] * vec[c],  r in [0,rows).  Lanes along rows, the
// column range split over thread groups whose partial sums meet in `part`.  Contains one
// __syncthreads when the split is active; the caller syncs before reading `out`.
__device__ __forceinline__ void row_dots(const double* __restrict__ M, int ld, int rows, int c0, int c1,
    const double* __restrict__ vec, double* __restrict__ out, double* __restrict__ part)
{
    const int tid = threadIdx.x, T = blockDim.x;
    const int rp = max(32, round32(rows));
    const int G = T / rp;
    if (G <= 1) {
        for (int r = tid; r < rows; r += T) {
            double s0 = 0.0, s1 = 0.0;
            int c = c0;
            for (; c + 1 < c1; c += 2) {
                s0 += M[r + size_t(c) * ld] * vec[c];
                s1 += M[r + size_t(c + 1) * ld] * vec[c + 1];
            }
            if (c < c1) s0 += M[r + size_t(c) * ld] * vec[c];
            out[r] = s0 + s1;
        }
        return;
    }
    const int g = tid / rp, r = tid - g * rp;
    if (g < G && r < rows) {
        double s = 0.0;
        for (int c = c0 + g; c < c1; c += G) s += M[r + size_t(c) * ld] * vec[c];
        part[g * rp + r] = s;
    }
    __syncthreads();
    if (tid < rows) {
        double s = part[tid];
        for (int k = 1; k < G; ++k) s += part[k * rp + tid];
        out[tid] = s;
    }
}

// "column dots": out[c] = sum_{r in [0,rows)} M[r + c*ld] * vec[r], c in [c0,c1).  One warp per column
// (lanes along the unit-stride rows, shuffle reduction), four columns in flight per warp.
__device__ __forceinline__ void col_dots(const double* __restrict__ M, int ld, int rows, int c0, int c1,
    const double* __restrict__ vec, double* __restrict__ out)
{
    const int lane = lane_id(), wp = warp_id(), nw = blockDim.x >> 5;
    for (int c = c0 + 4 * wp; c < c1; c += 4 * nw) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        const double* m0 = M + size_t(c) * ld;
        const bool h1 = c + 1 < c1, h2 = c + 2 < c1, h3 = c + 3 < c1;
        for (int r = lane; r < rows; r += 32) {
            const double vr = vec[r];
            s0 += m0[r] * vr;
            if (h1) s1 += m0[r + size_t(ld)] * vr;
            if (h2) s2 += m0[r + 2 * size_t(ld)] * vr;
            if (h3) s3 += m0[r + 3 * size_t(ld)] * vr;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            s3 += __shfl_xor_sync(0xffffffffu, s3, o);
        }
        if (lane == 0) {
            out[c] = s0;
            if (h1) out[c + 1] = s1;
            if (h2) out[c + 2] = s2;
            if (h3) out[c + 3] = s3;
        }
    }
}

// ---- dense factorisation helpers -----------------------------------------------------------------
// Upper Cholesky A = R'R in place (LINPACK dpofa's result, right-looking).  Only the upper triangle is
// read.  `row` is an n+1 scratch vector in shared memory.  Returns false (uniformly) if not PD.
__device__ inline bool chol_upper_inplace(double* __restrict__ J, int ld, int n, double* __restrict__ row)
{
    const int tid = threadIdx.x, T = blockDim.x;
    __syncthreads();
    double akk = J[0];
    for (int k = 0; k < n; ++k) {
        if (!(akk > 0.0)) return false;
        const double rkk = sqrt(akk);
        for (int j = k + 1 + tid; j < n; j += T) {
            const double v = J[k + size_t(j) * ld] / rkk;
            J[k + size_t(j) * ld] = v;
            row[j] = v;
        }
        __syncthreads();
        if (tid == 0) J[k + size_t(k) * ld] = rkk;
        // trailing update of the upper triangle: A[i,j] -= R[k,i] R[k,j], k < i <= j
        const int k1 = k + 1;
        tile_rc(n - k1, 0, n - k1, [&](int i, int j) {
            if (j >= i) J[(k1 + i) + size_t(k1 + j) * ld] -= row[k1 + i] * row[k1 + j];
        });
        __syncthreads();
        if (k1 < n) akk = J[k1 + size_t(k1) * ld];
    }
    return true;
}

// J := R^-1 for upper-triangular R, in place (LINPACK dpori's update order), strict lower triangle
// zeroed (qpgen2 label 21).  `row` (n+1) and `rowk` (n) are scratch vectors in shared memory.
__device__ inline void tri_inverse_upper_inplace(double* __restrict__ J, int ld, int n, double* __restrict__ row,
    double* __restrict__ rowk)
{
    const int tid = threadIdx.x, T = blockDim.x;
    for (int k = 0; k < n; ++k) {
        const double inv = 1.0 / J[k + size_t(k) * ld];
        for (int i = tid; i < k; i += T) row[i] = J[i + size_t(k) * ld] * (-inv);
        if (tid == 0) row[k] = inv;
        for (int j = k + 1 + tid; j < n; j += T) rowk[j] = J[k + size_t(j) * ld];
        __syncthreads();
        for (int i = tid; i <= k; i += T) J[i + size_t(k) * ld] = row[i];
        // columns j > k:  J[0..k-1, j] += t_j * row[0..k-1] ; J[k,j] = t_j * row[k]   (t_j = old J[k,j])
        tile_rc(k + 1, k + 1, n, [&](int i, int j) {
            const double t = rowk[j];
            if (i < k) J[i + size_t(j) * ld] += t * row[i];
            else J[k + size_t(j) * ld] = t * row[k];
        });
        __syncthreads();
    }
    for (int idx = tid; idx < n * n; idx += T) {
        const int i = idx % n, j = idx / n;
        if (i > j) J[i + size_t(j) * ld] = 0.0;
    }
    __syncthreads();
}

// ---- constraint access (dense form) -------------------------------------------------------------
// Stage the normal a_nvl of a general row into W.av (all threads), or describe a bound row.
__device__ __forceinline__ void gi_load_normal(const GiView& P, const GiWork& W, int nvl, int& bj, double& bsign)
{
    const int n = P.n, meq = P.meq, m = P.m, tid = threadIdx.x, T = blockDim.x;
    if (nvl < meq + m) {
        bj = -1; bsign = 0.0;
        if (nvl < meq) {
            const double sg = double(W.sgn[nvl]);
            if (W.A) for (int k = tid; k < n; k += T) W.av[k] = sg * W.A[nvl + size_t(k) * W.lda];
            else for (int k = tid; k < n; k += T) W.av[k] = sg * P.Aeq[nvl + size_t(k) * meq];
        } else {
            if (W.A) for (int k = tid; k < n; k += T) W.av[k] = -W.A[nvl + size_t(k) * W.lda];
            else for (int k = tid; k < n; k += T) W.av[k] = -P.Aineq[(nvl - meq) + size_t(k) * m];
        }
    } else {
        int j = nvl - meq - m;
        if (j < n) { bj = j; bsign = -1.0; }
        else { bj = j - n; bsign = 1.0; }
    }
}

// W.sl[i] = (general row i of [Aeq; Aineq]) . x   for all meq+m general rows
__device__ __forceinline__ void gi_general_products(const GiView& P, const GiWork& W)
{
    const int n = P.n, meq = P.meq, m = P.m;
    if (meq + m == 0) return;
    if (W.A) row_dots(W.A, W.lda, meq + m, 0, n, W.x, W.sl, W.part);
    else {
        if (meq) row_dots(P.Aeq, meq, meq, 0, n, W.x, W.sl, W.part);
        if (meq && m) __syncthreads();
        if (m) row_dots(P.Aineq, m, m, 0, n, W.x, W.sl + meq, W.part);
    }
}

// ---- the solver ---------------------------------------------------------------------------------
// All threads of the CTA call this with identical arguments.  Returns the QuadProg fail code
// (0 ok, 1 infeasible, 2 Hessian not PD; 3 = iteration cap, never seen in practice).
__device__ inline int gi_solve(const GiView& P, GiWork& W, const GiOut& O, double vsmall, int max_iter)
{
    const int n = P.n, meq = P.meq, m = P.m, mg = meq + m, q = mg + 2 * n;
    const int tid = threadIdx.x, T = blockDim.x, ld = W.ldj, lds = W.lds;
    double* __restrict__ J = W.J;
    double* __restrict__ S = W.S;
    double* scal = W.red + 2 * kMaxWarps; // small broadcast area

    // ---- 0. load --------------------------------------------------------------------------------
    for (int idx = tid; idx < n * n; idx += T) {
        const int i = idx % n, j = idx / n;
        J[i + size_t(j) * ld] = P.Q[idx];
    }
    if (W.A) {
        for (int idx = tid; idx < mg * n; idx += T) {
            const int i = idx % mg, k = idx / mg;
            W.A[i + size_t(k) * W.lda] = (i < meq) ? P.Aeq[i + size_t(k) * meq] : P.Aineq[(i - meq) + size_t(k) * m];
        }
    }
    for (int i = tid; i < n; i += T) {
        W.av[i] = -P.c[i];
        W.lb[i] = P.lb[i];
        W.ub[i] = P.ub[i];
        W.u[i] = 0.0;
        W.iact[i] = 0;
        W.rowmap[i] = i;
    }
    if (tid == 0) W.u[n] = 0.0;
    for (int i = tid; i < q; i += T) W.active[i] = 0;
    for (int i = tid; i < meq; i += T) W.sgn[i] = 1;
    __syncthreads();

    int fail = 0, nact = 0, iter0 = 0, iter1 = 0;

    // ---- 1. Cholesky Q = R'R (upper, in place) and 2. J = R^-1 in place ------------------------
#ifdef GC_PROFILE
    const long long gp0 = clock64();
#endif
    if (!gi_factor_blocked(J, ld, n, W.x)) fail = 2; // scratch: x d z r w v row sl part (gi_layout)
#ifdef GC_PROFILE
    const long long gp1 = clock64();
#endif
    if (fail == 0) {
        // ---- 3. unconstrained minimiser x = J J' (-c) ---------------------------------------------
        col_dots(J, ld, n, 0, n, W.av, W.d);
        __syncthreads();
        row_dots(J, ld, n, 0, n, W.d, W.x, W.part);
        // ---- 4. norms of the general rows ---------------------------------------------------------
        for (int i = tid; i < mg; i += T) {
            double s = 0.0;
            if (W.A) {
                for (int k = 0; k < n; ++k) { const double v = W.A[i + size_t(k) * W.lda]; s += v * v; }
            } else if (i < meq) {
                for (int k = 0; k < n; ++k) { const double v = P.Aeq[i + size_t(k) * meq]; s += v * v; }
            } else {
                for (int k = 0; k < n; ++k) { const double v = P.Aineq[(i - meq) + size_t(k) * m]; s += v * v; }
            }
            W.norm[i] = sqrt(s);
        }
        __syncthreads();

        // ---- 5. dual active-set iterations ---------------------------------------------------------
        for (;;) {
            ++iter0;
            if (iter0 > max_iter) { fail = 3; break; }
            // 5a. all slacks; most violated normalised constraint, lowest index on ties
            gi_general_products(P, W);
            __syncthreads();
            MinIdx best; best.v = 0.0; best.i = -1;
            double best_s = 0.0;
            for (int i = tid; i < q; i += T) {
                double s;
                if (i < meq) s = double(W.sgn[i]) * (W.sl[i] - P.beq[i]);
                else if (i < mg) s = P.bineq[i - meq] - W.sl[i];
                else s = gi_bound_slack(i - mg, n, mg, W.x, W.lb, W.ub, W.active);
                if (fabs(s) < vsmall) s = 0.0;
                if (i < meq) {
                    if (s > 0.0) W.sgn[i] = -W.sgn[i];
                    s = -fabs(s);
                }
                if (W.active[i]) s = 0.0;
                if (s < 0.0) {
                    const double nrm = (i < mg) ? W.norm[i] : 1.0;
                    MinIdx c; c.v = s / nrm; c.i = i;
                    const MinIdx nb = better(best, c);
                    if (nb.i != best.i) best_s = s;
                    best = nb;
                }
            }
            const MinIdx sel = block_argmin(best, W.red, W.redi);
            if (sel.i < 0) break; // optimal
            const int nvl = sel.i;
            if (best.i == nvl) scal[0] = best_s; // the owner publishes s_nvl
            __syncthreads();
            double s_nvl = scal[0];

            for (;;) { // label 55: step directions for constraint nvl
                int bj; double bsign;
                gi_load_normal(P, W, nvl, bj, bsign);
                __syncthreads();
                // d = J' a
                if (bj >= 0) {
                    for (int i = tid; i < n; i += T) W.d[i] = bsign * J[bj + size_t(i) * ld];
                } else col_dots(J, ld, n, 0, n, W.av, W.d);
                __syncthreads();
                // z = J2 d2 ; r = S d1 (rows of S through rowmap)
                row_dots(J, ld, n, nact, n, W.d, W.z, W.part);
                for (int i = tid; i < nact; i += T) {
                    const double* srow = S + W.rowmap[i];
                    double s0 = 0.0;
                    for (int k = 0; k < nact; ++k) s0 += srow[size_t(k) * lds] * W.d[k];
                    W.r[i] = s0;
                }
                __syncthreads();
                // t1 = min u_i/r_i over active inequalities with r_i > 0 ; z'z ; z'a ; d2'd2
                MinIdx tc; tc.v = 0.0; tc.i = -1;
                for (int i = tid; i < nact; i += T) {
                    if (W.iact[i] - 1 >= meq && W.r[i] > 0.0) {
                        MinIdx c; c.v = W.u[i] / W.r[i]; c.i = i;
                        tc = better(tc, c);
                    }
                }
                const MinIdx t1m = block_argmin(tc, W.red, W.redi);
                const bool t1inf = t1m.i < 0;
                const double t1 = t1m.v;
                const int it1 = t1m.i;
                double zz = 0.0, za = 0.0, dd = 0.0;
                for (int j = tid; j < n; j += T) {
                    const double zj = W.z[j];
                    zz += zj * zj;
                    if (bj < 0) za += zj * W.av[j];
                    if (j >= nact) { const double dj = W.d[j]; dd += dj * dj; }
                }
                block_sum2(zz, za, W.red);
                dd = block_sum(dd, W.red);
                if (bj >= 0) za = bsign * W.z[bj];

                bool do_drop = false;
                if (fabs(zz) <= vsmall) {
                    if (t1inf) { fail = 1; break; }
                    for (int i = tid; i < nact; i += T) W.u[i] -= t1 * W.r[i];
                    if (tid == 0) W.u[nact] += t1;
                    do_drop = true;
                } else {
                    double tt = -s_nvl / za;
                    bool t2min = true;
                    if (!t1inf && t1 < tt) { tt = t1; t2min = false; }
                    for (int j = tid; j < n; j += T) W.x[j] += tt * W.z[j];
                    for (int i = tid; i < nact; i += T) W.u[i] -= tt * W.r[i];
                    if (tid == 0) W.u[nact] += tt;
                    if (t2min) {
                        // ---- add constraint nvl: reflection H d2 = delta e1 applied to J2 ---------------
                        const double d0 = W.d[nact];
                        const double sigma = sqrt(dd);
                        const double delta = (d0 >= 0.0) ? -sigma : sigma;
                        const double tau = 1.0 / (sigma * (sigma + fabs(d0)));
                        __syncthreads(); // every thread has read d[nact]
                        // v = d2 - delta e1 (kept in d2), w = tau * J2 v = tau * (z - delta * J[:,nact])
                        for (int j = tid; j < n; j += T) W.w[j] = tau * (W.z[j] - delta * J[j + size_t(nact) * ld]);
                        if (tid == 0) W.d[nact] = d0 - delta;
                        __syncthreads();
                        tile_rc(n, nact, n, [&](int r_, int c_) { J[r_ + size_t(c_) * ld] -= W.w[r_] * W.d[c_]; });
                        // S gains the column [-r/delta ; 1/delta] and the row (0, .., 0, 1/delta)
                        const int newrow = W.rowmap[nact];
                        for (int i = tid; i < nact; i += T) {
                            S[W.rowmap[i] + size_t(nact) * lds] = -W.r[i] / delta;
                            S[newrow + size_t(i) * lds] = 0.0;
                        }
                        if (tid == 0) {
                            S[newrow + size_t(nact) * lds] = 1.0 / delta;
                            W.iact[nact] = nvl + 1;
                            W.active[nvl] = 1;
                        }
                        ++nact;
                        __syncthreads();
                        break; // back to 5a
                    } else {
                        // partial step: refresh s_nvl at the new x (with the equality sign rule)
                        __syncthreads(); // x complete
                        double s;
                        if (bj >= 0) {
                            s = (bsign < 0.0) ? W.ub[bj] - W.x[bj] : W.x[bj] - W.lb[bj];
                        } else {
                            double acc = 0.0; // av holds the signed normal, so a'x = av . x
                            for (int k = tid; k < n; k += T) acc += W.av[k] * W.x[k];
                            acc = block_sum(acc, W.red);
                            if (nvl < meq) s = acc - double(W.sgn[nvl]) * P.beq[nvl];
                            else s = acc + P.bineq[nvl - meq];
                        }
                        if (nvl < meq) {
                            __syncthreads(); // all threads have read sgn[nvl]
                            if (s > 0.0 && tid == 0) W.sgn[nvl] = -W.sgn[nvl];
                            s = -fabs(s);
                        }
                        s_nvl = s;
                        do_drop = true;
                    }
                }
                if (do_drop) {
                    // ---- drop the it1-th active constraint: reflection with last column ~ row p of S ----
                    __syncthreads(); // u updates visible
                    const int p = it1;
                    const int dropped = (tid == 0) ? W.iact[p] - 1 : 0; // used by thread 0 only, which also clears iact below
                    const int prow = W.rowmap[p];
                    if (nact > 1) {
                        // v = row p of S (nact entries); rho = |v|; w = v - gamma e_last; d := tau * w
                        double vv = 0.0;
                        for (int k = tid; k < nact; k += T) { const double t_ = S[prow + size_t(k) * lds]; W.v[k] = t_; vv += t_ * t_; }
                        vv = block_sum(vv, W.red);
                        const double rho = sqrt(vv);
                        const double vl = W.v[nact - 1];
                        const double gamma = (vl >= 0.0) ? -rho : rho;
                        const double tau = 1.0 / (rho * (rho + fabs(vl)));
                        __syncthreads(); // every thread has read v[nact-1]
                        if (tid == 0) W.v[nact - 1] = vl - gamma;
                        __syncthreads();
                        for (int k = tid; k < nact; k += T) W.d[k] = tau * W.v[k];
                        // J1 w (all rows of J) and S w (active rows), then the two rank-1 updates
                        row_dots(J, ld, n, 0, nact, W.v, W.w, W.part);
                        for (int i = tid; i < nact; i += T) {
                            const double* srow = S + W.rowmap[i];
                            double s0 = 0.0;
                            for (int k = 0; k < nact; ++k) s0 += srow[size_t(k) * lds] * W.v[k];
                            W.r[i] = s0; // r is recomputed at label 55
                        }
                        __syncthreads();
                        tile_rc(n, 0, nact, [&](int r_, int c_) { J[r_ + size_t(c_) * ld] -= W.w[r_] * W.d[c_]; });
                        for (int i = tid; i < nact; i += T) {
                            if (i == p) continue;
                            double* srow = S + W.rowmap[i];
                            const double ri = W.r[i];
                            for (int k = 0; k < nact - 1; ++k) srow[size_t(k) * lds] -= ri * W.d[k];
                        }
                        __syncthreads();
                        // close the gap at position p in u / iact / rowmap (one warp, register staged)
                        if (warp_id() == 0) {
                            const int lane = lane_id();
                            for (int base = p; base < nact - 1; base += 32) {
                                const int k = base + lane;
                                double uu = 0.0; int ia = 0, rm = 0;
                                if (k < nact - 1) { uu = W.u[k + 1]; ia = W.iact[k + 1]; rm = W.rowmap[k + 1]; }
                                __syncwarp();
                                if (k < nact - 1) { W.u[k] = uu; W.iact[k] = ia; W.rowmap[k] = rm; }
                                __syncwarp();
                            }
                            if (lane == 0) W.rowmap[nact - 1] = prow;
                        }
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
    // ---- 6. results -------------------------------------------------------------------------------
#ifdef GC_PROFILE
    if (tid == 0 && blockIdx.x == 0) printf("GIPROF n=%d iters=%d factor=%lld rest=%lld\n", n, iter0, gp1 - gp0, clock64() - gp1);
#endif
    if (O.x) for (int i = tid; i < n; i += T) O.x[i] = (fail == 2) ? 0.0 : W.x[i];
    if (O.iact) for (int i = tid; i < n; i += T) O.iact[i] = (i < nact) ? W.iact[i] : 0;
    if (tid == 0) {
        if (O.status) *O.status = fail;
        if (O.iters) { O.iters[0] = iter0; O.iters[1] = iter1; }
        if (O.nact) *O.nact = nact;
    }
    return fail;
}

} // namespace cb
