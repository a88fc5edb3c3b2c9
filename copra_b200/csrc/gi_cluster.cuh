// gi_cluster.cuh -- K5+K6 when the n x n factor J does not fit one SM: a thread-block CLUSTER of C
// CTAs (C = 2, 4 or 8) owns one QP instance.  J = L^-T Q is distributed by ROWS over the C shared
// memories (CTA c keeps rows [c*nr, (c+1)*nr) of every column), so that
//   * d = J'a       : each CTA forms partial column dots over its row slab; the partials are reduce-
//                     scattered and the sums all-gathered through distributed shared memory (DSMEM stores
//                     straight into the peers' buffers, two cluster barriers);
//   * z = J2 d2     : row dots, purely local; the slabs of z are all-gathered through DSMEM;
//   * ADD / DROP    : the Householder rank-1 updates of gi_solver.cuh, purely local on the slab;
//   * slacks        : the general rows are dealt out to the CTAs (each needs the full x, which every CTA
//                     keeps), the per-CTA arg-min candidates meet through DSMEM.
// The explicit inverse factor S = R^-1 of the active set (nact x nact, touched lightly) lives in an
// L2-resident global workspace: every CTA computes r = S d1 redundantly, rank 0 alone updates S.
// Factorisation (Cholesky, inverse, unconstrained minimiser) is done beforehand by k5_factor_kernel,
// one CTA per instance, and handed over through global memory.
// Three cluster barriers per active-set iteration.  Same iterates as gi_solver.cuh / qpgen2.
#pragma once
#ifdef GC_PROFILE
#include <cstdio>
#endif
#include "common.cuh"
#include "engine.cuh"
#include "gi_solver.cuh"

#include <cooperative_groups.h>

namespace cb {

namespace cg = cooperative_groups;

constexpr int kClMaxC = 8;

struct GcLayout {
    int n, C, nr, ld, seg, mshare, threads;
    int nb;          // pivots per block step of the distributed factorisation (panel rows that fit the scratch region)
    size_t scratch;  // doubles from oX that are free while the factorisation runs (x d z v r a w t recv part)
    size_t oJ, oX, oD, oZ, oV, oR, oU, oA, oW, oT, oLb, oUb, oRecv, oNorm, oPart, oRed, oCand; // doubles
    size_t oIact, oRowmap, oRedI, oCtl, oActive, oSgn, bytes;                        // bytes
};

__host__ __device__ inline GcLayout gc_layout(int n, int meq, int m, int C, int threads)
{
    GcLayout L;
    L.n = n; L.C = C; L.threads = threads;
    L.nr = (n + C - 1) / C;           // rows of J per CTA
    L.ld = odd_ld(L.nr);              // odd: conflict-free along rows and along columns
    L.seg = (n + C - 1) / C;          // columns per CTA in the reduce-scatter of d
    L.mshare = (meq + m + C - 1) / C; // general rows per CTA
    size_t o = 0;
    auto take = [&](size_t cnt) { size_t at = o; o += (cnt + 1) & ~size_t(1); return at; };
    L.oJ = take(size_t(L.ld) * n);
    // iteration vectors that hold nothing while the factorisation runs come first and contiguously: the factorisation
    // aliases them as its broadcast panel (nb pivot rows) and per-row coefficient table
    L.oX = take(n); L.oD = take(n); L.oZ = take(size_t(L.nr) * C); L.oV = take(n); L.oR = take(n);
    L.oA = take(L.nr); L.oW = take(L.nr); L.oT = take(n);
    L.oRecv = take(size_t(L.seg) * C);
    L.oPart = take(2 * size_t(threads) + 64);
    L.scratch = o - L.oX;
    {
        const size_t per = size_t(n) + 2 + size_t(L.nr) + 2; // panel row + coefficient column + reciprocal, per pivot
        const size_t fit = L.scratch / per;
        L.nb = int(fit < 1 ? 1 : (fit > 8 ? 8 : fit));
    }
    L.oU = take(n + 2); L.oLb = take(n); L.oUb = take(n);
    L.oNorm = take(L.mshare);
    L.oRed = take(4 * kMaxWarps);
    L.oCand = take(8 * kClMaxC); // per-rank slots: [3r..3r+2] arg-min candidate, [3C+2r..] partial scalars
    size_t b = o * sizeof(double);
    L.oIact = b; b += sizeof(int) * size_t(n);
    L.oRowmap = b; b += sizeof(int) * size_t(n);
    L.oRedI = b; b += sizeof(int) * kMaxWarps;
    L.oCtl = b; b += sizeof(int) * 8;
    L.oActive = b; b += size_t(meq + m + 2 * n);
    L.oSgn = b; b += size_t(meq > 0 ? meq : 1);
    L.bytes = (b + 15) & ~size_t(15);
    return L;
}

struct GcWork {
    double *J, *x, *d, *z, *v, *r, *u, *a, *w, *t, *lb, *ub, *recv, *norm, *part, *red, *cand;
    int *iact, *rowmap, *redi, *ctl;
    unsigned char* active;
    signed char* sgn;
};

__device__ inline GcWork gc_carve(const GcLayout& L, unsigned char* smem)
{
    GcWork W;
    double* b = reinterpret_cast<double*>(smem);
    W.J = b + L.oJ; W.x = b + L.oX; W.d = b + L.oD; W.z = b + L.oZ; W.v = b + L.oV; W.r = b + L.oR; W.u = b + L.oU;
    W.a = b + L.oA; W.w = b + L.oW; W.t = b + L.oT; W.lb = b + L.oLb; W.ub = b + L.oUb; W.recv = b + L.oRecv; W.norm = b + L.oNorm; W.part = b + L.oPart;
    W.red = b + L.oRed; W.cand = b + L.oCand;
    W.iact = reinterpret_cast<int*>(smem + L.oIact);
    W.rowmap = reinterpret_cast<int*>(smem + L.oRowmap);
    W.redi = reinterpret_cast<int*>(smem + L.oRedI);
    W.ctl = reinterpret_cast<int*>(smem + L.oCtl);
    W.active = smem + L.oActive;
    W.sgn = reinterpret_cast<signed char*>(smem + L.oSgn);
    return W;
}

// All threads of all CTAs of the cluster call this with identical arguments.
//   S  : global workspace of lds x n doubles for this cluster
__device__ inline int gc_solve(const GiView& P, const GcLayout& L, GcWork& W, double* S, int lds, const GiOut& O, double vsmall,
    int max_iter)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = int(cluster.block_rank()), C = L.C;
    const int n = L.n, meq = P.meq, m = P.m, mg = meq + m, q = mg + 2 * n, ld = L.ld, seg = L.seg;
    const int tid = threadIdx.x, T = blockDim.x;
    const int r0 = min(n, rank * L.nr), nrc = max(0, min(L.nr, n - r0));       // my rows of J: [r0, r0+nrc)
    const int g0 = min(mg, rank * L.mshare), g1 = min(mg, g0 + L.mshare);     // my general rows
    const int bshare = (2 * n + C - 1) / C, b0 = min(2 * n, rank * bshare), b1 = min(2 * n, b0 + bshare); // my bound rows
    const int c0 = min(n, rank * seg), myseg = max(0, min(seg, n - c0));       // my columns of d in the reduce-scatter
    double* __restrict__ J = W.J;
    double* scal = W.red + 2 * kMaxWarps;

#ifdef GC_PROFILE
    long long tp0 = clock64(), tp1 = 0, tp2 = 0, tp3 = 0, ta = 0, tb = 0, tg = 0, td = 0, tmark = 0;
#endif
    // ---- load my slab of Q --------------------------------------------------------------------------
    for (int idx = tid; idx < nrc * n; idx += T) {
        const int r = idx % nrc, c = idx / nrc;
        J[r + size_t(c) * ld] = P.Q[(r0 + r) + size_t(c) * n];
    }
    for (int i = tid; i < n; i += T) {
        W.lb[i] = P.lb[i];
        W.ub[i] = P.ub[i];
        W.u[i] = 0.0;
        W.iact[i] = 0;
        W.rowmap[i] = i;
    }
    if (tid < 2) W.u[n + tid] = 0.0;
    for (int i = tid; i < q; i += T) W.active[i] = 0;
    for (int i = tid; i < meq; i += T) W.sgn[i] = 1;
    for (int i = g0 + tid; i < g1; i += T) {
        const double* ap = (i < meq) ? P.Aeq + i : P.Aineq + (i - meq);
        const size_t as = (i < meq) ? meq : m;
        double s = 0.0;
        for (int k = 0; k < n; ++k) { const double v = ap[size_t(k) * as]; s += v * v; }
        W.norm[i - g0] = sqrt(s);
    }
    __syncthreads();

#ifdef GC_PROFILE
    tp1 = clock64();
#endif
    // ---- distributed, blocked factorisation: Q = R'R (LINPACK dpofa) and J = R^-1 (dpori) in ONE pass ----------
    // Per pivot p the fused sweep is  J[i,j] = (i == p ? 0 : J[i,j]) + mult_p[j] * coef_p[i]  (j > p, i <= j) with
    // mult_p = row p of R and coef_p[i] = -J[i,p]/R[p,p] (i < p), 1/R[p,p] (i == p), -R[p,i] (i > p): every entry sees
    // the same updates in the same order as the two LINPACK loops.  Pivots are taken nb at a time (a block never
    // straddles two slabs): the owner factors its nb rows locally (block barriers only), broadcasts them as one panel
    // (the peers pull it through DSMEM), and every CTA applies a rank-nb update to its slab -- two cluster barriers
    // per nb pivots instead of two per pivot.
    bool pd = true;
    {
        const int NB = L.nb, pstride = n + 2;
        double* panel = W.x;                                   // NB x (n+2): [pp][j] = R[p,j] (j >= p), [pp][n] = pivot ok
        double* coefb = panel + size_t(NB) * pstride;          // nr x NB   : coef_p[i] for my rows
        double* invb = coefb + size_t(L.nr) * NB;              // NB        : 1 / R[p,p]
#ifdef GC_PROFILE
        long long f1 = 0, f2 = 0, f3 = 0, f4 = 0, f5 = 0, fm = clock64();
#define GC_TICK(acc) { const long long t_ = clock64(); acc += t_ - fm; fm = t_; }
#else
#define GC_TICK(acc)
#endif
        for (int k0 = 0; k0 < n;) {
            const int owner = k0 / L.nr;
            const int k1 = min(min(k0 + NB, n), (owner + 1) * L.nr), nb = k1 - k0;
            if (rank == owner) {
                bool ok = true;
                for (int p = k0; p < k1; ++p) {
                    const int pp = p - k0, lr = p - r0;
                    double* prow = panel + size_t(pp) * pstride;
                    const double akk = J[lr + size_t(p) * ld];
                    ok = ok && akk > 0.0;
                    const double rkk = ok ? sqrt(akk) : 1.0;
                    __syncthreads(); // everyone has read the pivot before the thread that owns it overwrites it
                    for (int j = p + tid; j < n; j += T) {
                        const double v = (j == p) ? rkk : J[lr + size_t(j) * ld] / rkk;
                        J[lr + size_t(j) * ld] = v;
                        prow[j] = v;
                    }
                    if (tid == 0) prow[n] = ok ? 1.0 : 0.0;
                    __syncthreads();
                    // the rest of my block rows: A[i,j] -= R[p,i] R[p,j], p < i < k1, j >= i.  One thread per column (the slab's
                    // leading dimension is odd, so a warp walking along a row is conflict-free), rows in a short loop.
                    const int rows = k1 - 1 - p;
                    if (rows > 0) {
                        for (int c_ = p + 1 + tid; c_ < n; c_ += T) {
                            const double mc = prow[c_];
                            double* col = J + lr + 1 + size_t(c_) * ld;
                            const int rmax = min(rows, c_ - p); // rows i = p+1+r_ with i <= c_
                            for (int r_ = 0; r_ < rmax; ++r_) col[r_] = fma(mc, -prow[p + 1 + r_], col[r_]);
                        }
                        __syncthreads();
                    }
                }
            }
            GC_TICK(f1)
            cluster.sync();
            GC_TICK(f2)
            if (rank != owner) {
                // PULL the panel (columns >= k0 and the flags) from the owner's shared memory: the copy is issued by all the
                // peers' SMs in parallel (an owner-side push of nb x n x (C-1) remote stores was issue-bound on one SM)
                const double* src = cluster.map_shared_rank(panel, owner);
                const int width = pstride - k0;
                for (int idx = tid; idx < nb * width; idx += T) {
                    const int pp = idx / width, j = k0 + (idx - pp * width);
                    panel[size_t(pp) * pstride + j] = src[size_t(pp) * pstride + j];
                }
                __syncthreads();
            }
            for (int pp = 0; pp < nb; ++pp) pd = pd && panel[size_t(pp) * pstride + n] != 0.0;
            if (!pd) break;
            GC_TICK(f3)
            if (tid < nb) invb[tid] = 1.0 / panel[size_t(tid) * pstride + k0 + tid];
            __syncthreads();
            // coefficients of my rows, pivot by pivot (a short serial fold per row), and the finished columns k0..k1-1
            for (int r = tid; r < nrc; r += T) {
                const int i = r0 + r;
                for (int pp = 0; pp < nb; ++pp) {
                    const int p = k0 + pp;
                    double c;
                    if (i > p) c = -panel[size_t(pp) * pstride + i];
                    else if (i == p) c = invb[pp];
                    else {
                        // J[i,p] as the sweeps of the earlier pivots of this block leave it (those with q < i were already
                        // applied by the owner's local update when i is one of the block rows)
                        double a = J[r + size_t(p) * ld];
                        for (int qq = max(0, i - k0); qq < pp; ++qq)
                            a = fma(panel[size_t(qq) * pstride + p], coefb[r + size_t(qq) * L.nr], (i == k0 + qq) ? 0.0 : a);
                        c = a * (-invb[pp]);
                    }
                    coefb[r + size_t(pp) * L.nr] = c;
                    if (i <= p) J[r + size_t(p) * ld] = c;
                }
            }
            __syncthreads();
            GC_TICK(f4)
            // rank-nb update of the columns to the right of the block.  With full 8-pivot panels it is a (rows x 8) x (8 x cols)
            // product on the FP64 tensor cores (DMMA.8x8x4, two k-steps per 8x8 tile of my slab; tiles dealt out to the
            // warps); entries below the diagonal are updated too -- they are never read and are zeroed at the end.  The
            // owner's 8 block rows follow the per-pivot rule (pivot p zeroes row p, earlier pivots are already applied),
            // so they take the scalar path, one thread per column.
            if (NB == 8) {
                const int lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
                const int MT = (nrc + 7) >> 3, NT = (n - k1 + 7) >> 3;
                const int skip = (rank == owner) ? (k0 - r0) >> 3 : -1; // blocks are 8-aligned inside a slab
                for (int t = warp; t < MT * NT; t += nwarp) {
                    const int mt = t % MT, nt = t / MT;
                    if (mt == skip) continue;
                    const int ar = (mt << 3) + (lane >> 2), ak = lane & 3;
                    const double a0 = (ar < nrc && ak < nb) ? coefb[ar + size_t(ak) * L.nr] : 0.0;
                    const double a1 = (ar < nrc && ak + 4 < nb) ? coefb[ar + size_t(ak + 4) * L.nr] : 0.0;
                    const int cb_ = k1 + (nt << 3), bc = min(cb_ + (lane >> 2), n - 1);
                    const double b0 = panel[size_t(ak) * pstride + bc], b1 = panel[size_t(ak + 4) * pstride + bc];
                    const int cc = cb_ + 2 * (lane & 3);
                    const bool v0 = ar < nrc && cc < n, v1 = ar < nrc && cc + 1 < n;
                    double c0_ = v0 ? J[ar + size_t(cc) * ld] : 0.0, c1_ = v1 ? J[ar + size_t(cc + 1) * ld] : 0.0;
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0_), "+d"(c1_) : "d"(a0), "d"(b0));
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0_), "+d"(c1_) : "d"(a1), "d"(b1));
                    if (v0) J[ar + size_t(cc) * ld] = c0_;
                    if (v1) J[ar + size_t(cc + 1) * ld] = c1_;
                }
                if (rank == owner) {
                    for (int c_ = k1 + tid; c_ < n; c_ += T) {
                        for (int rr = 0; rr < nb; ++rr) {
                            const int r_ = k0 - r0 + rr;
                            double a = 0.0; // pivot k0+rr zeroes its own row first
                            for (int pp = rr; pp < nb; ++pp) a = fma(panel[size_t(pp) * pstride + c_], coefb[r_ + size_t(pp) * L.nr], a);
                            J[r_ + size_t(c_) * ld] = a;
                        }
                    }
                }
            } else {
                // scalar fallback (panel narrower than 8 pivots): lanes along my rows, the column range dealt out to thread
                // groups; a thread's nb coefficients stay in registers for its whole column walk
                const int rp = max(32, round32(nrc));
                const int G = max(1, T / rp), g = tid / rp, r_ = tid - g * rp;
                if (g < G && r_ < nrc) {
                    const int i = r0 + r_;
                    const int pp0 = (i >= k0 && i < k1) ? i - k0 : 0; // block rows: pivots before i were applied by the owner
                    double cf[8];
#pragma unroll
                    for (int pp = 0; pp < 8; ++pp) cf[pp] = (pp < nb) ? coefb[r_ + size_t(pp) * L.nr] : 0.0;
                    for (int c_ = max(k1, i) + ((g - max(k1, i)) % G + G) % G; c_ < n; c_ += G) {
                        double a = J[r_ + size_t(c_) * ld];
#pragma unroll
                        for (int pp = 0; pp < 8; ++pp)
                            if (pp >= pp0 && pp < nb) a = fma(panel[size_t(pp) * pstride + c_], cf[pp], (i == k0 + pp) ? 0.0 : a);
                        J[r_ + size_t(c_) * ld] = a;
                    }
                }
            }
            cluster.sync(); // everyone is done with this panel before the next owner overwrites it
            GC_TICK(f5)
            k0 = k1;
        }
#ifdef GC_PROFILE
        if (rank == 0 && tid == 0) printf("GCFACT own=%lld sync1=%lld pull=%lld coef=%lld upd+sync2=%lld\n", f1, f2, f3, f4, f5);
#endif
    }
    if (!pd) {
        if (rank == 0) {
            for (int i = tid; i < n; i += T) { if (O.x) O.x[i] = 0.0; if (O.iact) O.iact[i] = 0; }
            if (tid == 0) {
                if (O.status) *O.status = 2;
                if (O.iters) { O.iters[0] = 0; O.iters[1] = 0; }
                if (O.nact) *O.nact = 0;
            }
        }
        return 2;
    }
    for (int idx = tid; idx < nrc * n; idx += T) { // strict lower triangle := 0 (qpgen2 label 21)
        const int r = idx % nrc, c = idx / nrc;
        if (r0 + r > c) J[r + size_t(c) * ld] = 0.0;
    }
    __syncthreads();
#ifdef GC_PROFILE
    tp2 = clock64();
#endif
    // ---- unconstrained minimiser x = J (J' (-c)): same exchange pattern as the iteration ------------------
    {
        for (int r = tid; r < nrc; r += T) W.a[r] = -P.c[r0 + r];
        __syncthreads();
        for (int c = tid; c < n; c += T) {
            const double* col = J + size_t(c) * ld;
            double p0 = 0.0;
            for (int r = 0; r < nrc; ++r) p0 += col[r] * W.a[r];
            const int owner = c / seg;
            cluster.map_shared_rank(W.recv, owner)[rank * seg + (c - owner * seg)] = p0;
        }
        cluster.sync();
        for (int cl = tid; cl < myseg; cl += T) {
            double sum = W.recv[cl];
            for (int k = 1; k < C; ++k) sum += W.recv[k * seg + cl];
            for (int t = 0; t < C; ++t) cluster.map_shared_rank(W.d, t)[c0 + cl] = sum;
        }
        cluster.sync();
        row_dots(J, ld, nrc, 0, n, W.d, W.w, W.part);
        __syncthreads();
        for (int r = tid; r < nrc; r += T) {
            const double xr = W.w[r];
            for (int t = 0; t < C; ++t) cluster.map_shared_rank(W.x, t)[r0 + r] = xr;
        }
        cluster.sync();
    }

#ifdef GC_PROFILE
    tp3 = clock64();
#endif
    int fail = 0, nact = 0, iter0 = 0, iter1 = 0;
    bool pending = false;
    int pc0 = 0;
    for (;;) {
#ifdef GC_PROFILE
        tmark = clock64();
#endif
        ++iter0;
        if (iter0 > max_iter) { fail = 3; break; }
        // ================= alpha: pending rank-1 on my slab; slacks of my rows; candidate exchange ======
        if (pending) {
            tile_rc(nrc, pc0, n, [&](int r_, int c_) { J[r_ + size_t(c_) * ld] -= W.w[r_] * W.v[c_]; });
            pending = false;
        }
        MinIdx best; best.v = 0.0; best.i = -1;
        double best_s = 0.0;
        auto consider = [&](int i, double s, double nrm) {
            if (fabs(s) < vsmall) s = 0.0;
            if (i < meq) {
                if (s > 0.0) W.sgn[i] = -W.sgn[i];
                s = -fabs(s);
            }
            if (W.active[i]) s = 0.0;
            if (s < 0.0) {
                MinIdx c; c.v = s / nrm; c.i = i;
                const MinIdx nb = better(best, c);
                if (nb.i != best.i) best_s = s;
                best = nb;
            }
        };
        for (int base = g0; base < g1; base += T) { // general rows: rows along lanes, columns split over thread groups
            const int cnt = min(T, g1 - base), rp = round32(cnt), G = max(1, T / rp);
            const int g = tid / rp, r = tid - g * rp;
            if (g < G && r < cnt) {
                const int i = base + r;
                const double* ap = (i < meq) ? P.Aeq + i : P.Aineq + (i - meq);
                const size_t as = (i < meq) ? meq : m;
                double s = 0.0;
                int k = g;
                for (; k + 7 * G < n; k += 8 * G) { // 8 independent L2 loads in flight per thread
                    double av[8];
#pragma unroll
                    for (int u_ = 0; u_ < 8; ++u_) av[u_] = ap[size_t(k + u_ * G) * as];
#pragma unroll
                    for (int u_ = 0; u_ < 8; ++u_) s += av[u_] * W.x[k + u_ * G];
                }
                for (; k < n; k += G) s += ap[size_t(k) * as] * W.x[k];
                W.part[g * rp + r] = s;
            }
            __syncthreads();
            if (tid < cnt) {
                const int i = base + tid;
                double p = W.part[tid];
                for (int k = 1; k < G; ++k) p += W.part[k * rp + tid];
                const double s = (i < meq) ? double(W.sgn[i]) * (p - P.beq[i]) : P.bineq[i - meq] - p;
                consider(i, s, W.norm[i - g0]);
            }
            __syncthreads();
        }
        for (int j = b0 + tid; j < b1; j += T) // bound rows: upper (-x_j >= -ub_j) then lower (x_j >= lb_j)
            consider(mg + j, gi_bound_slack(j, n, mg, W.x, W.lb, W.ub, W.active), 1.0);
        {
            const MinIdx mine = block_argmin(best, W.red, W.redi);
            if (best.i == mine.i && mine.i >= 0) { scal[0] = best_s; scal[1] = (mine.i < meq) ? double(W.sgn[mine.i]) : -1.0; }
            __syncthreads();
            if (tid < C) { // publish my candidate into slot `rank` of CTA `tid`
                double* rc = cluster.map_shared_rank(W.cand, tid);
                rc[4 * rank + 0] = mine.v;
                rc[4 * rank + 1] = double(mine.i);
                rc[4 * rank + 2] = (mine.i >= 0) ? scal[0] : 0.0;
                rc[4 * rank + 3] = (mine.i >= 0) ? scal[1] : 0.0;
            }
        }
        cluster.sync();
#ifdef GC_PROFILE
        { const long long t = clock64(); ta += t - tmark; tmark = t; }
#endif
        // ================= beta: select; d = J' a through reduce-scatter + all-gather ====================
        MinIdx sel; sel.v = 0.0; sel.i = -1;
        double s_nvl = 0.0, asign = -1.0;
        for (int k = 0; k < C; ++k) {
            MinIdx c; c.v = W.cand[4 * k]; c.i = int(W.cand[4 * k + 1]);
            const MinIdx nb = better(sel, c);
            if (nb.i != sel.i) { s_nvl = W.cand[4 * k + 2]; asign = W.cand[4 * k + 3]; }
            sel = nb;
        }
        if (sel.i < 0) break; // optimal (every CTA takes the same branch)
        const int nvl = sel.i;
        int bj = -1;
        if (nvl >= mg) { bj = nvl - mg; asign = -1.0; if (bj >= n) { bj -= n; asign = 1.0; } }

        for (;;) { // label 55
            if (bj >= 0) {
                // d = asign * (row bj of J): its owner writes it into every CTA's d
                const int owner = bj / L.nr;
                if (rank == owner) {
                    for (int c = tid; c < n; c += T) {
                        const double v = asign * J[(bj - r0) + size_t(c) * ld];
                        for (int t = 0; t < C; ++t) cluster.map_shared_rank(W.d, t)[c] = v;
                    }
                }
                for (int r = tid; r < nrc; r += T) W.a[r] = (r0 + r == bj) ? asign : 0.0;
                cluster.sync();
            } else {
                const double* ap = (nvl < meq) ? P.Aeq + nvl : P.Aineq + (nvl - meq);
                const size_t as = (nvl < meq) ? meq : m;
                for (int r = tid; r < nrc; r += T) W.a[r] = asign * ap[size_t(r0 + r) * as];
                __syncthreads();
                for (int c = tid; c < n; c += T) { // partial column dots over my slab
                    const double* col = J + size_t(c) * ld;
                    double p0 = 0.0, p1 = 0.0;
                    int r = 0;
                    for (; r + 1 < nrc; r += 2) { p0 += col[r] * W.a[r]; p1 += col[r + 1] * W.a[r + 1]; }
                    if (r < nrc) p0 += col[r] * W.a[r];
                    const int owner = c / seg;
                    cluster.map_shared_rank(W.recv, owner)[rank * seg + (c - owner * seg)] = p0 + p1;
                }
                cluster.sync();
                for (int cl = tid; cl < myseg; cl += T) {
                    double sum = W.recv[cl];
                    for (int k = 1; k < C; ++k) sum += W.recv[k * seg + cl];
                    for (int t = 0; t < C; ++t) cluster.map_shared_rank(W.d, t)[c0 + cl] = sum;
                }
                cluster.sync();
            }
#ifdef GC_PROFILE
            { const long long t = clock64(); tb += t - tmark; tmark = t; }
#endif
            // ============= gamma: z slab (all-gathered), r = S d1, candidates, norms =======================
            row_dots(J, ld, nrc, nact, n, W.d, W.w, W.part); // w temporarily holds my slab of z
            __syncthreads();
            {   // my share of r = S d1: active rows [a0,a1), rows along lanes, columns split over thread groups
                const int ashare = (nact + C - 1) / C, a0 = min(nact, rank * ashare), a1 = min(nact, a0 + ashare);
                const int cnt = a1 - a0, rp = max(32, round32(cnt)), G = max(1, T / rp);
                const int g = tid / rp, rr = tid - g * rp;
                if (g < G && rr < cnt) {
                    const double* srow = S + W.rowmap[a0 + rr];
                    double s0 = 0.0;
                    int k = g;
                    for (; k + 3 * G < nact; k += 4 * G) {
                        const double v0 = __ldcg(srow + size_t(k) * lds), v1 = __ldcg(srow + size_t(k + G) * lds);
                        const double v2 = __ldcg(srow + size_t(k + 2 * G) * lds), v3 = __ldcg(srow + size_t(k + 3 * G) * lds);
                        s0 += v0 * W.d[k] + v1 * W.d[k + G] + v2 * W.d[k + 2 * G] + v3 * W.d[k + 3 * G];
                    }
                    for (; k < nact; k += G) s0 += __ldcg(srow + size_t(k) * lds) * W.d[k];
                    W.part[g * rp + rr] = s0;
                }
                __syncthreads();
                if (tid < cnt) {
                    double s0 = W.part[tid];
                    for (int k = 1; k < G; ++k) s0 += W.part[k * rp + tid];
                    for (int t = 0; t < C; ++t) cluster.map_shared_rank(W.r, t)[a0 + tid] = s0;
                }
            }
            double zz = 0.0, za = 0.0, dd = 0.0;
            for (int r = tid; r < nrc; r += T) {
                const double zr = W.w[r];
                zz += zr * zr;
                za += zr * W.a[r];
                for (int t = 0; t < C; ++t) cluster.map_shared_rank(W.z, t)[r0 + r] = zr;
            }
            for (int c = nact + tid; c < n; c += T) { const double dc = W.d[c]; dd += dc * dc; }
            block_sum2(zz, za, W.red);
            dd = block_sum(dd, W.red);
            if (tid < C) {
                double* rc = cluster.map_shared_rank(W.cand, tid);
                rc[4 * C + 2 * rank + 0] = zz;
                rc[4 * C + 2 * rank + 1] = za;
            }
            cluster.sync();
#ifdef GC_PROFILE
            { const long long t = clock64(); tg += t - tmark; tmark = t; }
#endif
            // ============= delta: step lengths, x / u, reflection vectors ==================================
            zz = 0.0; za = 0.0;
            for (int k = 0; k < C; ++k) { zz += W.cand[4 * C + 2 * k]; za += W.cand[4 * C + 2 * k + 1]; }
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
                    // ---- ADD: H d2 = delta e1 ; my slab of w = tau (z - delta J[:,nact]) ; v = d2 - delta e1 ----
                    const double d0 = W.d[nact];
                    const double sigma = sqrt(dd);
                    const double delta = (d0 >= 0.0) ? -sigma : sigma;
                    const double tau = 1.0 / (sigma * (sigma + fabs(d0)));
                    __syncthreads(); // row_dots result in w consumed by everyone above
                    for (int r = tid; r < nrc; r += T) W.w[r] = tau * (W.z[r0 + r] - delta * J[r + size_t(nact) * ld]);
                    for (int c = nact + tid; c < n; c += T) W.v[c] = (c == nact) ? d0 - delta : W.d[c];
                    const int newrow = W.rowmap[nact];
                    if (rank == 0) { // rank 0 alone maintains S; the next cluster barrier publishes it
                        for (int i = tid; i < nact; i += T) {
                            S[W.rowmap[i] + size_t(nact) * lds] = -W.r[i] / delta;
                            S[newrow + size_t(i) * lds] = 0.0;
                        }
                        if (tid == 0) S[newrow + size_t(nact) * lds] = 1.0 / delta;
                        __threadfence();
                    }
                    if (tid == 0) {
                        W.iact[nact] = nvl + 1;
                        W.active[nvl] = 1;
                    }
                    pending = true;
                    pc0 = nact;
                    ++nact;
                    __syncthreads();
                    break; // -> alpha
                } else {
                    // partial step: refresh s_nvl at the new x (every CTA redundantly; equality sign rule)
                    __syncthreads();
                    double s;
                    if (bj >= 0) s = (asign < 0.0) ? W.ub[bj] - W.x[bj] : W.x[bj] - W.lb[bj];
                    else {
                        const double* ap = (nvl < meq) ? P.Aeq + nvl : P.Aineq + (nvl - meq);
                        const size_t as = (nvl < meq) ? meq : m;
                        double acc = 0.0;
                        for (int k = tid; k < n; k += T) acc += ap[size_t(k) * as] * W.x[k];
                        acc = block_sum(acc, W.red);
                        s = (nvl < meq) ? asign * (acc - P.beq[nvl]) : P.bineq[nvl - meq] - acc;
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
                // ---- DROP the it1-th active constraint -------------------------------------------------------
                __syncthreads();
                const int p = it1;
                const int dropped = (tid == 0) ? W.iact[p] - 1 : 0; // used by thread 0 only, which also clears iact below
                const int prow = W.rowmap[p];
                if (nact > 1) {
                    // w = (row p of S) - gamma e_last ; t = tau w   (every CTA redundantly; row p is not modified below)
                    double vv = 0.0;
                    for (int k = tid; k < nact; k += T) { const double t_ = __ldcg(S + prow + size_t(k) * lds); W.v[k] = t_; vv += t_ * t_; }
                    vv = block_sum(vv, W.red);
                    const double rho = sqrt(vv);
                    const double vl = W.v[nact - 1];
                    const double gamma = (vl >= 0.0) ? -rho : rho;
                    const double tau = 1.0 / (rho * (rho + fabs(vl)));
                    __syncthreads();
                    if (tid == 0) W.v[nact - 1] = vl - gamma;
                    __syncthreads();
                    for (int k = tid; k < nact; k += T) W.t[k] = tau * W.v[k];
                    row_dots(J, ld, nrc, 0, nact, W.v, W.w, W.part); // my slab of J1 w
                    __syncthreads();
                    tile_rc(nrc, 0, nact, [&](int r_, int c_) { J[r_ + size_t(c_) * ld] -= W.w[r_] * W.t[c_]; });
                    if (rank == 0) {
                        for (int i = tid; i < nact; i += T) {
                            if (i == p) continue;
                            double* srow = S + W.rowmap[i];
                            double sw = 0.0;
                            for (int k = 0; k < nact; ++k) sw += __ldcg(srow + size_t(k) * lds) * W.v[k];
                            for (int k = 0; k < nact - 1; ++k) srow[size_t(k) * lds] = __ldcg(srow + size_t(k) * lds) - sw * W.t[k];
                        }
                        __threadfence();
                    }
                    // close the gap at position p in u / iact / rowmap
                    __syncthreads();
                    for (int base = p; base < nact - 1; base += T) {
                        const int k = base + tid;
                        double uu = 0.0; int ia = 0, rm = 0;
                        const bool mv = k < nact - 1;
                        if (mv) { uu = W.u[k + 1]; ia = W.iact[k + 1]; rm = W.rowmap[k + 1]; }
                        __syncthreads();
                        if (mv) { W.u[k] = uu; W.iact[k] = ia; W.rowmap[k] = rm; }
                        __syncthreads();
                    }
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
                cluster.sync(); // peers may still read d / z of this pass; S updates become visible
                continue;       // label 55
            }
        }
        if (fail != 0) break;
    }
#ifdef GC_PROFILE
    td = clock64() - tp3 - ta - tb - tg;
    if (rank == 0 && tid == 0)
        printf("GCPROF n=%d iters=%d load=%lld chol+inv=%lld x0=%lld alpha=%lld beta=%lld gamma=%lld delta+rest=%lld total=%lld\n", n, iter0, tp1 - tp0, tp2 - tp1,
            tp3 - tp2, ta, tb, tg, td, clock64() - tp0);
#endif
    __syncthreads();
    if (rank == 0) {
        if (O.x) for (int i = tid; i < n; i += T) O.x[i] = W.x[i];
        if (O.iact) for (int i = tid; i < n; i += T) O.iact[i] = (i < nact) ? W.iact[i] : 0;
        if (tid == 0) {
            if (O.status) *O.status = fail;
            if (O.iters) { O.iters[0] = iter0; O.iters[1] = iter1; }
            if (O.nact) *O.nact = nact;
        }
    }
    return fail;
}

} // namespace cb
