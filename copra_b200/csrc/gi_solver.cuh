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
// B200-first formulation (same iterates as qpgen2 in exact arithmetic):
//   * J = R^-1 (n x n) lives in shared memory (column-major, odd leading dimension => conflict-free
//     for both thread-per-row and thread-per-column sweeps) or, when it does not fit one SM, in an
//     L2-resident global workspace.
//   * qpgen2 keeps the triangular factor R of the active set and back-substitutes r = R^-1 d1 (a
//     length-nact serial chain).  Here the INVERSE factor S = R^-1 is kept instead: r = S d1 is a
//     fully parallel triangular mat-vec, an added constraint appends the column [-r/delta; 1/delta],
//     and a dropped constraint is a chain of plane rotations whose coefficients come from a prefix
//     norm scan of one row of S (no serial Hessenberg sweep).
//   * the Givens chain that folds d2 into its first component after an add is likewise computed
//     from a suffix-norm scan, then applied thread-per-row of J with one load + one store per entry.
#pragma once
#include "common.cuh"
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
    int n, meq, m, ldj, lds;
    int j_smem, s_smem, a_smem;
    // offsets in doubles
    size_t oJ, oS, oA, oX, oD, oZ, oAv, oR, oU, oGc, oGs, oNu, oRow, oNorm, oLb, oUb, oRed;
    // then ints / bytes (byte offsets from the start)
    size_t oIact, oRowmap, oRedI, oActive, oSgn, bytes;
};

__host__ __device__ inline GiLayout gi_layout(int n, int meq, int m, int j_smem, int s_smem, int a_smem)
{
    GiLayout L;
    L.n = n; L.meq = meq; L.m = m;
    L.ldj = odd_ld(n); L.lds = odd_ld(n);
    L.j_smem = j_smem; L.s_smem = s_smem; L.a_smem = a_smem;
    size_t o = 0;
    L.oJ = o; if (j_smem) o += size_t(L.ldj) * n;
    L.oS = o; if (s_smem) o += size_t(L.lds) * n;
    L.oA = o; if (a_smem) o += size_t(meq + m) * n;
    L.oX = o; o += n;
    L.oD = o; o += n;
    L.oZ = o; o += n;
    L.oAv = o; o += n;
    L.oR = o; o += n;
    L.oU = o; o += n + 1;
    L.oGc = o; o += n;
    L.oGs = o; o += n;
    L.oNu = o; o += n;
    L.oRow = o; o += n + 1;
    L.oNorm = o; o += meq + m;
    L.oLb = o; o += n;
    L.oUb = o; o += n;
    L.oRed = o; o += 2 * kMaxWarps;
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
    double *J, *S, *A; // A: cached [Aeq; Aineq] rows, (meq+m) x n col-major (or nullptr)
    double *x, *d, *z, *av, *r, *u, *gc, *gs, *nu, *row, *norm, *lb, *ub, *red;
    int *iact, *rowmap, *redi;
    unsigned char* active;
    signed char* sgn;
    int ldj, lds;
};

__device__ inline GiWork gi_carve(const GiLayout& L, unsigned char* smem, double* gJ, double* gS)
{
    GiWork W;
    double* base = reinterpret_cast<double*>(smem);
    W.J = L.j_smem ? base + L.oJ : gJ;
    W.S = L.s_smem ? base + L.oS : gS;
    W.A = L.a_smem ? base + L.oA : nullptr;
    W.x = base + L.oX; W.d = base + L.oD; W.z = base + L.oZ; W.av = base + L.oAv; W.r = base + L.oR;
    W.u = base + L.oU; W.gc = base + L.oGc; W.gs = base + L.oGs; W.nu = base + L.oNu; W.row = base + L.oRow;
    W.norm = base + L.oNorm; W.lb = base + L.oLb; W.ub = base + L.oUb; W.red = base + L.oRed;
    W.iact = reinterpret_cast<int*>(smem + L.oIact);
    W.rowmap = reinterpret_cast<int*>(smem + L.oRowmap);
    W.redi = reinterpret_cast<int*>(smem + L.oRedI);
    W.active = smem + L.oActive;
    W.sgn = reinterpret_cast<signed char*>(smem + L.oSgn);
    W.ldj = L.ldj; W.lds = L.lds;
    return W;
}

// ---- dense factorisation helpers (all threads of the CTA call them with identical arguments) ----
// Upper Cholesky A = R'R in place (result of LINPACK dpofa; right-looking so that every entry sees
// its updates in the same k-ascending order).  Only the upper triangle is read.  `row` is an n+1
// scratch vector in shared memory.  Returns false (uniformly) if A is not positive definite.
__device__ inline bool chol_upper_inplace(double* __restrict__ J, int ld, int n, double* __restrict__ row)
{
    const int tid = threadIdx.x, T = blockDim.x;
    __syncthreads();
    double akk = J[0];
    for (int k = 0; k < n; ++k) {
        if (!(akk > 0.0)) return false;
        const double rkk = sqrt(akk);
        for (int j = k + 1 + tid; j < n; j += T) {
            double v = J[k + size_t(j) * ld] / rkk;
            J[k + size_t(j) * ld] = v;
            row[j] = v;
        }
        __syncthreads();
        if (tid == 0) J[k + size_t(k) * ld] = rkk;
        for (int j = k + 1 + tid; j < n; j += T) {
            const double rkj = row[j];
            double* col = J + size_t(j) * ld;
            for (int i = k + 1; i <= j; ++i) col[i] -= row[i] * rkj;
        }
        __syncthreads();
        if (k + 1 < n) akk = J[(k + 1) + size_t(k + 1) * ld];
    }
    return true;
}

// J := R^-1 for upper-triangular R, in place (LINPACK dpori's update order, parallel over columns),
// then the strict lower triangle is zeroed (qpgen2 label 21).
__device__ inline void tri_inverse_upper_inplace(double* __restrict__ J, int ld, int n, double* __restrict__ row)
{
    const int tid = threadIdx.x, T = blockDim.x;
    for (int k = 0; k < n; ++k) {
        const double inv = 1.0 / J[k + size_t(k) * ld];
        for (int i = tid; i < k; i += T) row[i] = J[i + size_t(k) * ld] * (-inv);
        if (tid == 0) row[k] = inv;
        __syncthreads();
        for (int i = tid; i <= k; i += T) J[i + size_t(k) * ld] = row[i];
        for (int j = k + 1 + tid; j < n; j += T) {
            double* col = J + size_t(j) * ld;
            const double t = col[k];
            for (int i = 0; i < k; ++i) col[i] += t * row[i];
            col[k] = t * row[k];
        }
        __syncthreads();
    }
    for (int j = tid; j < n; j += T) {
        double* col = J + size_t(j) * ld;
        for (int i = j + 1; i < n; ++i) col[i] = 0.0;
    }
    __syncthreads();
}

// ---- constraint access (dense form) -------------------------------------------------------------
// slack of constraint i at the current x:  a_i'x - b_i  in quadprog's ">= 0 is feasible" convention.
__device__ __forceinline__ double gi_slack(const GiView& P, const GiWork& W, int i)
{
    const int n = P.n, meq = P.meq, m = P.m;
    if (i < meq + m) {
        double s = 0.0;
        if (W.A) {
            const double* a = W.A + i;
            const int ld = meq + m;
            for (int k = 0; k < n; ++k) s += a[size_t(k) * ld] * W.x[k];
        } else if (i < meq) {
            const double* a = P.Aeq + i;
            for (int k = 0; k < n; ++k) s += a[size_t(k) * meq] * W.x[k];
        } else {
            const double* a = P.Aineq + (i - meq);
            for (int k = 0; k < n; ++k) s += a[size_t(k) * m] * W.x[k];
        }
        if (i < meq) return double(W.sgn[i]) * (s - P.beq[i]); // sgn*(Aeq x) - sgn*beq
        return P.bineq[i - meq] - s;                           // (-Aineq x) - (-bineq)
    }
    int j = i - meq - m;
    if (j < n) return W.ub[j] - W.x[j]; // -x_j - (-ub_j)
    j -= n;
    return W.x[j] - W.lb[j];            //  x_j - lb_j
}

// Stage the normal a_nvl of a dense row into W.av (all threads), or return the bound descriptor.
// returns sign (+1/-1) and column j through `bj` for bound rows (bj >= 0), else bj = -1.
__device__ __forceinline__ void gi_load_normal(const GiView& P, const GiWork& W, int nvl, int& bj, double& bsign)
{
    const int n = P.n, meq = P.meq, m = P.m, tid = threadIdx.x, T = blockDim.x;
    if (nvl < meq + m) {
        bj = -1; bsign = 0.0;
        if (nvl < meq) {
            const double sg = double(W.sgn[nvl]);
            if (W.A) for (int k = tid; k < n; k += T) W.av[k] = sg * W.A[nvl + size_t(k) * (meq + m)];
            else for (int k = tid; k < n; k += T) W.av[k] = sg * P.Aeq[nvl + size_t(k) * meq];
        } else {
            if (W.A) for (int k = tid; k < n; k += T) W.av[k] = -W.A[nvl + size_t(k) * (meq + m)];
            else for (int k = tid; k < n; k += T) W.av[k] = -P.Aineq[(nvl - meq) + size_t(k) * m];
        }
    } else {
        int j = nvl - meq - m;
        if (j < n) { bj = j; bsign = -1.0; }
        else { bj = j - n; bsign = 1.0; }
    }
}

// ---- the solver ---------------------------------------------------------------------------------
// All threads of the CTA call this with identical arguments.  Returns the QuadProg fail code.
__device__ inline int gi_solve(const GiView& P, GiWork& W, const GiOut& O, double vsmall, int max_iter)
{
    const int n = P.n, meq = P.meq, m = P.m, q = meq + m + 2 * n;
    const int tid = threadIdx.x, T = blockDim.x, ld = W.ldj, lds = W.lds;
    double* __restrict__ J = W.J;
    double* __restrict__ S = W.S;

    // ---- 0. load --------------------------------------------------------------------------------
    for (int idx = tid; idx < n * n; idx += T) {
        int i = idx % n, j = idx / n;
        J[i + size_t(j) * ld] = P.Q[idx];
    }
    if (W.A) {
        const int ma = meq + m;
        for (int idx = tid; idx < ma * n; idx += T) {
            int i = idx % ma, k = idx / ma;
            W.A[idx] = (i < meq) ? P.Aeq[i + size_t(k) * meq] : P.Aineq[(i - meq) + size_t(k) * m];
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
    if (!chol_upper_inplace(J, ld, n, W.row)) fail = 2;
    if (fail == 0) {
        tri_inverse_upper_inplace(J, ld, n, W.row);
        // ---- 3. unconstrained minimiser x = J J' (-c) ---------------------------------------------
        for (int i = tid; i < n; i += T) {
            const double* col = J + size_t(i) * ld;
            double s = 0.0;
            for (int j = 0; j <= i; ++j) s += col[j] * W.av[j];
            W.d[i] = s;
        }
        __syncthreads();
        for (int j = tid; j < n; j += T) {
            double s = 0.0;
            for (int i = j; i < n; ++i) s += J[j + size_t(i) * ld] * W.d[i];
            W.x[j] = s;
        }
        // ---- 4. constraint norms ------------------------------------------------------------------
        for (int i = tid; i < meq + m; i += T) {
            double s = 0.0;
            if (W.A) {
                for (int k = 0; k < n; ++k) { double v = W.A[i + size_t(k) * (meq + m)]; s += v * v; }
            } else if (i < meq) {
                for (int k = 0; k < n; ++k) { double v = P.Aeq[i + size_t(k) * meq]; s += v * v; }
            } else {
                for (int k = 0; k < n; ++k) { double v = P.Aineq[(i - meq) + size_t(k) * m]; s += v * v; }
            }
            W.norm[i] = sqrt(s);
        }
        __syncthreads();

        // ---- 5. dual active-set iterations ---------------------------------------------------------
        for (;;) {
            ++iter0;
            if (iter0 > max_iter) { fail = 3; break; }
            // 5a. all slacks, most violated normalised constraint (lowest index on ties)
            MinIdx best; best.v = 0.0; best.i = -1;
            double best_s = 0.0;
            for (int i = tid; i < q; i += T) {
                double s = gi_slack(P, W, i);
                if (fabs(s) < vsmall) s = 0.0;
                if (i < meq) {
                    if (s > 0.0) W.sgn[i] = -W.sgn[i];
                    s = -fabs(s);
                }
                if (W.active[i]) s = 0.0;
                if (s < 0.0) {
                    const double nrm = (i < meq + m) ? W.norm[i] : 1.0;
                    MinIdx c; c.v = s / nrm; c.i = i;
                    MinIdx nb = better(best, c);
                    if (nb.i != best.i) best_s = s;
                    best = nb;
                }
            }
            const MinIdx sel = block_argmin(best, W.red, W.redi);
            if (sel.i < 0) break; // optimal
            const int nvl = sel.i;
            if (best.i == nvl) W.red[2 * kMaxWarps - 1] = best_s; // owner publishes s_nvl
            __syncthreads();
            double s_nvl = W.red[2 * kMaxWarps - 1];
            __syncthreads();

            for (;;) { // label 55: (re)compute the step directions for constraint nvl
                int bj; double bsign;
                gi_load_normal(P, W, nvl, bj, bsign);
                __syncthreads();
                // d = J' a
                if (bj >= 0) {
                    for (int i = tid; i < n; i += T) W.d[i] = bsign * J[bj + size_t(i) * ld];
                } else {
                    for (int i = tid; i < n; i += T) {
                        const double* col = J + size_t(i) * ld;
                        double s0 = 0.0;
                        for (int k = 0; k < n; ++k) s0 += col[k] * W.av[k];
                        W.d[i] = s0;
                    }
                }
                __syncthreads();
                // z = J2 d2 ; r = S d1
                for (int j = tid; j < n; j += T) {
                    double s0 = 0.0;
                    for (int i = nact; i < n; ++i) s0 += J[j + size_t(i) * ld] * W.d[i];
                    W.z[j] = s0;
                }
                for (int i = tid; i < nact; i += T) {
                    const double* srow = S + W.rowmap[i];
                    double s0 = 0.0;
                    for (int k = i; k < nact; ++k) s0 += srow[size_t(k) * lds] * W.d[k];
                    W.r[i] = s0;
                }
                __syncthreads();
                // t1 = min u_i/r_i over active inequalities with r_i > 0 ; z'z ; z'a
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
                double zz = 0.0, za = 0.0;
                for (int j = tid; j < n; j += T) {
                    const double zj = W.z[j];
                    zz += zj * zj;
                    if (bj < 0) za += zj * W.av[j];
                }
                block_sum2(zz, za, W.red);
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
                        // ---- add constraint nvl ------------------------------------------------
                        const int L = n - nact;
                        double delta;
                        if (L == 1) {
                            delta = W.d[n - 1];
                        } else {
                            // suffix scans over v = d[nact..n): S_i = sum_{k>=i} v_k^2 ; sg_i = sign of first non-zero v_k, k>=i
                            // done by warp 0: chunk per lane, shuffle scan of chunk totals.
                            if (warp_id() == 0) {
                                const int lane = lane_id();
                                const int chunk = (L + 31) >> 5;
                                const int lo = lane * chunk, hi = min(L, lo + chunk);
                                double tot = 0.0; int sg = 0;
                                for (int i = hi - 1; i >= lo; --i) {
                                    const double v = W.d[nact + i];
                                    tot += v * v;
                                    if (v != 0.0) sg = v > 0.0 ? 1 : -1;
                                }
                                // exclusive suffix over lanes (lanes > me)
                                double suf = 0.0; int ssg = 0;
                                for (int src = 31; src >= 0; --src) {
                                    const double t_ = __shfl_sync(0xffffffffu, tot, src);
                                    const int g_ = __shfl_sync(0xffffffffu, sg, src);
                                    if (src > lane) { suf += t_; if (g_ != 0) ssg = g_; }
                                }
                                // walk own chunk from the back producing S_i (in gs) and sg_i (in nu as +-1/0)
                                double run = suf; int rsg = ssg;
                                for (int i = hi - 1; i >= lo; --i) {
                                    const double v = W.d[nact + i];
                                    run += v * v;
                                    if (v != 0.0) rsg = v > 0.0 ? 1 : -1;
                                    W.gs[nact + i] = run;
                                    W.nu[nact + i] = double(rsg);
                                }
                            }
                            __syncthreads();
                            // rotation for the pair (i-1, i), i = 1..L-1 (local index), stored at [nact+i]
                            // gc==1 => skip, gc==0 => swap (qpgen2 labels 160-180)
                            for (int i = 1 + tid; i < L; i += T) {
                                const double Si = W.gs[nact + i], Sim1 = W.gs[nact + i - 1];
                                const double vprev = W.d[nact + i - 1];
                                double gc = 1.0, gsv = 0.0, nuv = 0.0; // default: skip
                                if (Si != 0.0) {
                                    const double cur = (i == L - 1) ? W.d[nact + i] : W.nu[nact + i] * sqrt(Si);
                                    double temp = sqrt(Sim1);
                                    if (vprev < 0.0) temp = -temp;
                                    gc = vprev / temp;
                                    gsv = cur / temp;
                                    if (gc != 1.0 && gc != 0.0) nuv = gsv / (1.0 + gc);
                                }
                                W.gc[nact + i] = gc;
                                W.row[nact + i] = gsv;
                                W.av[nact + i] = nuv; // av is free between label-55 passes once d is formed
                            }
                            __syncthreads();
                            // apply to J, thread per row, from the last pair down
                            for (int j = tid; j < n; j += T) {
                                double* rowp = J + j;
                                double hi = rowp[size_t(n - 1) * ld];
                                for (int i = L - 1; i >= 1; --i) {
                                    const int col = nact + i;
                                    const double gc = W.gc[col];
                                    double lo = rowp[size_t(col - 1) * ld];
                                    if (gc == 1.0) {
                                        rowp[size_t(col) * ld] = hi;
                                        hi = lo;
                                    } else if (gc == 0.0) {
                                        rowp[size_t(col) * ld] = lo;
                                        // hi (old column `col`) moves to column col-1
                                    } else {
                                        const double gsv = W.row[col], nuv = W.av[col];
                                        const double temp = gc * lo + gsv * hi;
                                        rowp[size_t(col) * ld] = nuv * (lo + temp) - hi;
                                        hi = temp;
                                    }
                                }
                                rowp[size_t(nact) * ld] = hi;
                            }
                            // delta = value left in d[nact] by the chain
                            {
                                const double S1 = W.gs[nact + 1], S0 = W.gs[nact];
                                const double v0 = W.d[nact];
                                const double gc1 = W.gc[nact + 1];
                                if (S1 == 0.0 || gc1 == 1.0) delta = v0;
                                else if (gc1 == 0.0) delta = (L - 1 == 1) ? W.d[nact + 1] : W.nu[nact + 1] * sqrt(S1);
                                else delta = (v0 < 0.0) ? -sqrt(S0) : sqrt(S0);
                            }
                        }
                        // S gets the column [-r/delta ; 1/delta]
                        for (int i = tid; i < nact; i += T) S[W.rowmap[i] + size_t(nact) * lds] = -W.r[i] / delta;
                        if (tid == 0) {
                            S[W.rowmap[nact] + size_t(nact) * lds] = 1.0 / delta;
                            W.iact[nact] = nvl + 1;
                            W.active[nvl] = 1;
                        }
                        ++nact;
                        __syncthreads();
                        break; // back to 5a
                    } else {
                        // partial step: refresh s_nvl at the new x (with the equality sign rule)
                        __syncthreads(); // x complete
                        if (tid == 0) {
                            double s = gi_slack(P, W, nvl);
                            if (nvl < meq) {
                                if (s > 0.0) W.sgn[nvl] = -W.sgn[nvl];
                                s = -fabs(s);
                            }
                            W.red[2 * kMaxWarps - 1] = s;
                        }
                        __syncthreads();
                        s_nvl = W.red[2 * kMaxWarps - 1];
                        do_drop = true;
                    }
                }
                if (do_drop) {
                    // ---- drop the it1-th active constraint -----------------------------------------
                    __syncthreads(); // u updates visible
                    const int p = it1;
                    const int dropped = W.iact[p] - 1;
                    if (p < nact - 1) {
                        // v = row p of S (columns p..nact-1); prefix norms by warp 0
                        const double* srow = S + W.rowmap[p];
                        const int Lr = nact - p;
                        if (warp_id() == 0) {
                            const int lane = lane_id();
                            const int chunk = (Lr + 31) >> 5;
                            const int lo = lane * chunk, hi = min(Lr, lo + chunk);
                            double tot = 0.0;
                            for (int i = lo; i < hi; ++i) { const double v = srow[size_t(p + i) * lds]; tot += v * v; }
                            double pre = 0.0;
                            for (int src = 0; src < 32; ++src) {
                                const double t_ = __shfl_sync(0xffffffffu, tot, src);
                                if (src < lane) pre += t_;
                            }
                            double run = pre;
                            for (int i = lo; i < hi; ++i) {
                                const double v = srow[size_t(p + i) * lds];
                                run += v * v;
                                W.gs[p + i] = run;    // P_k
                                W.nu[p + i] = v;      // v_k
                            }
                        }
                        __syncthreads();
                        // plane (k,k+1), k = p..nact-2: coefficients stored at [k]
                        for (int k = p + tid; k < nact - 1; k += T) {
                            const double b = W.nu[k + 1];
                            const double h = sqrt(W.gs[k + 1]);
                            const double temp = (b >= 0.0) ? -h : h;
                            const double a = (k == p) ? W.nu[p] : ((W.nu[k] >= 0.0) ? -sqrt(W.gs[k]) : sqrt(W.gs[k]));
                            W.gc[k] = -b / temp;
                            W.row[k] = a / temp;
                        }
                        __syncthreads();
                        // rotate columns of J (all n rows) ...
                        for (int j = tid; j < n; j += T) {
                            double* rowp = J + j;
                            double a = rowp[size_t(p) * ld];
                            for (int k = p; k < nact - 1; ++k) {
                                const double gc = W.gc[k], gsv = W.row[k];
                                const double b = rowp[size_t(k + 1) * ld];
                                rowp[size_t(k) * ld] = gc * a + gsv * b;
                                a = gsv * a - gc * b;
                            }
                            rowp[size_t(nact - 1) * ld] = a;
                        }
                        // ... and of S (logical rows i != p; row i > p starts at column i-1)
                        for (int i = tid; i < nact; i += T) {
                            if (i == p) continue;
                            double* sr = S + W.rowmap[i];
                            const int k0 = (i > p) ? max(p, i - 1) : p;
                            double a = (k0 >= i) ? sr[size_t(k0) * lds] : 0.0;
                            for (int k = k0; k < nact - 1; ++k) {
                                const double gc = W.gc[k], gsv = W.row[k];
                                const double b = sr[size_t(k + 1) * lds];
                                sr[size_t(k) * lds] = gc * a + gsv * b;
                                a = gsv * a - gc * b;
                            }
                            // the last column is discarded
                        }
                        __syncthreads();
                        // shift u, iact, rowmap down over position p (single warp, registers)
                        if (warp_id() == 0) {
                            const int lane = lane_id();
                            const int freed = W.rowmap[p];
                            for (int base = p; base < nact - 1; base += 32) {
                                const int k = base + lane;
                                double uu = 0.0; int ia = 0, rm = 0;
                                if (k < nact - 1) { uu = W.u[k + 1]; ia = W.iact[k + 1]; rm = W.rowmap[k + 1]; }
                                __syncwarp();
                                if (k < nact - 1) { W.u[k] = uu; W.iact[k] = ia; W.rowmap[k] = rm; }
                                __syncwarp();
                            }
                            if (lane == 0) W.rowmap[nact - 1] = freed;
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
