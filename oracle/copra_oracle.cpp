/*
 * copra_oracle.cpp -- CPU ORACLE: test infrastructure, NOT the product.  PARITY PINNING: the
 * reference cannot be compiled offline (no Eigen / eigen-quadprog / gfortran), and
 * its own tests hold no golden vectors for this path, so this restatement is pinned
 * by (i) the known-answer vector KA-1 of the reference's `Problem` fixture (tests/systems.h:11-38;
 * tests/golden/ka_problem.json), (ii) an independent numpy restatement of K1-K5
 * (oracle/py/copra_numpy.py), (iii) solver-independent KKT checks.  Bit-level agreement of
 * tie-breaks with eigen-quadprog's Fortran remains "parity unpinned".
 *
 * Each function cites the reference file:line (relative to /root/reference) it restates.
 * Scalar loops, column-major storage, no Eigen.  Nothing here is used by libcopra_b200.so.
 */
#include "copra_oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;

struct Mat { // column-major, ld == rows
    int r = 0, c = 0;
    std::vector<double> a;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), a(size_t(r_) * c_, 0.0) {}
    Mat(int r_, int c_, const double* src) : r(r_), c(c_), a(src, src + size_t(r_) * c_) {}
    double& operator()(int i, int j) { return a[size_t(j) * r + i]; }
    double operator()(int i, int j) const { return a[size_t(j) * r + i]; }
    void resize(int r_, int c_) { r = r_; c = c_; a.assign(size_t(r_) * c_, 0.0); }
    void setZero() { std::fill(a.begin(), a.end(), 0.0); }
};
typedef std::vector<double> Vec;

const double kInf = std::numeric_limits<double>::infinity();

[[noreturn]] void domainError(const std::string& m) { throw std::domain_error(m); }
[[noreturn]] void runtimeError(const std::string& m) { throw std::runtime_error(m); }

// C(rc x cc) = A.block(ar,ac,rc,k) * B.block(br,bc,k,cc)
void mulBlock(const Mat& A, int ar, int ac, const Mat& B, int br, int bc, int rc, int k, int cc, Mat& C)
{
    C.resize(rc, cc);
    for (int j = 0; j < cc; ++j)
        for (int l = 0; l < k; ++l) {
            double b = B(br + l, bc + j);
            for (int i = 0; i < rc; ++i) C(i, j) += A(ar + i, ac + l) * b;
        }
}

/* ---------------------------------------------------------------------------------------------
 * AutoSpan -- reference src/AutoSpan.cpp:10-47
 * ------------------------------------------------------------------------------------------- */
void spanMatrix(Mat& mat, int new_dim, int addCols = 0)
{
    int matRows = mat.r;
    if (new_dim == matRows) return; // AutoSpan.cpp:13-15
    int matCols = mat.c;
    Mat tmp = mat;
    int nrStep = matRows ? new_dim / matRows : 0;
    if (nrStep * matRows != new_dim) domainError("spanMatrix: bad new_dim"); // :20-22
    mat.resize(new_dim, matCols * (nrStep + addCols));
    for (int i = 0; i < nrStep; ++i)
        for (int cc = 0; cc < matCols; ++cc)
            for (int rr = 0; rr < matRows; ++rr) mat(i * matRows + rr, i * matCols + cc) = tmp(rr, cc);
}
void spanVector(Vec& vec, int new_dim)
{
    int vecRows = int(vec.size());
    if (new_dim == vecRows) return; // AutoSpan.cpp:32-34
    int nrStep = vecRows ? new_dim / vecRows : 0;
    Vec tmp = vec;
    if (nrStep * vecRows != new_dim) domainError("spanVector: bad new_dim"); // :39-41
    vec.resize(new_dim);
    for (int i = 1; i < nrStep; ++i)
        for (int k = 0; k < vecRows; ++k) vec[i * vecRows + k] = tmp[k];
}

/* ---------------------------------------------------------------------------------------------
 * PreviewSystem -- reference src/PreviewSystem.cpp:16-74
 * ------------------------------------------------------------------------------------------- */
struct PreviewSystem {
    int nrUStep = 0, nrXStep = 0, xDim = 0, uDim = 0, fullXDim = 0, fullUDim = 0;
    Vec x0, d, xi;
    Mat A, B, Phi, Psi;

    void system(int nx, int nu, int N, const double* A_, const double* B_, const double* d_, const double* x0_)
    {
        if (N <= 0) domainError("The number of step sould be a positive number! "); // :31-33
        nrUStep = N; nrXStep = N + 1; xDim = nx; uDim = nu;
        fullXDim = xDim * nrXStep; fullUDim = uDim * nrUStep;
        x0.assign(x0_, x0_ + nx);
        A = Mat(nx, nx, A_); B = Mat(nx, nu, B_); d.assign(d_, d_ + nx);
        Phi.resize(fullXDim, xDim); Psi.resize(fullXDim, fullUDim); xi.assign(fullXDim, 0.0);
        for (int i = 0; i < xDim; ++i) Phi(i, i) = 1.0; // :51-52
    }
    void updateSystem() // :57-74
    {
        for (int j = 0; j < xDim; ++j)
            for (int i = 0; i < xDim; ++i) Phi(xDim + i, j) = A(i, j);
        for (int j = 0; j < uDim; ++j)
            for (int i = 0; i < xDim; ++i) Psi(xDim + i, j) = B(i, j);
        for (int i = 0; i < xDim; ++i) xi[xDim + i] = d[i];
        for (int i = 2; i < nrXStep; ++i) {
            // Phi_i = A * Phi_{i-1}
            for (int j = 0; j < xDim; ++j)
                for (int rr = 0; rr < xDim; ++rr) {
                    double s = 0;
                    for (int k = 0; k < xDim; ++k) s += A(rr, k) * Phi((i - 1) * xDim + k, j);
                    Phi(i * xDim + rr, j) = s;
                }
            // Psi_{i,0} = A * Psi_{i-1,0}
            for (int j = 0; j < uDim; ++j)
                for (int rr = 0; rr < xDim; ++rr) {
                    double s = 0;
                    for (int k = 0; k < xDim; ++k) s += A(rr, k) * Psi((i - 1) * xDim + k, j);
                    Psi(i * xDim + rr, j) = s;
                }
            // Psi_{i,j} = Psi_{i-1,j-1}
            for (int j = 1; j < i; ++j)
                for (int cc = 0; cc < uDim; ++cc)
                    for (int rr = 0; rr < xDim; ++rr)
                        Psi(i * xDim + rr, j * uDim + cc) = Psi((i - 1) * xDim + rr, (j - 1) * uDim + cc);
            // xi_i = A * xi_{i-1} + d
            for (int rr = 0; rr < xDim; ++rr) {
                double s = 0;
                for (int k = 0; k < xDim; ++k) s += A(rr, k) * xi[(i - 1) * xDim + k];
                xi[i * xDim + rr] = s + d[rr];
            }
        }
    }
};

/* ---------------------------------------------------------------------------------------------
 * Cost functions -- reference src/costFunctions.cpp:14-215, include/costFunctions.h:54-67
 * ------------------------------------------------------------------------------------------- */
struct Cost {
    int kind;
    bool fullSizeEntry = false;
    Mat M, N, Q, E;
    Vec p, weights, c, f;

    explicit Cost(const orc_cost& s) : kind(s.kind)
    {
        if (kind != ORC_COST_CONTROL) M = Mat(s.rows, s.colsM, s.M);
        if (kind == ORC_COST_CONTROL || kind == ORC_COST_MIXED) N = Mat(s.rows, s.colsN, s.N);
        p.assign(s.p, s.p + s.rows);
        weights.assign(s.rows, 1.0); // ctor: weights_ = Ones(p.rows()) (costFunctions.h:117)
        if (s.w) setWeights(s.w, s.wrows);
        if (s.autospan) autoSpan();
    }
    void setWeights(const double* w, int wrows) // costFunctions.h:54-67
    {
        int have = int(weights.size());
        if (wrows == have) weights.assign(w, w + wrows);
        else if (wrows > 0 && have % wrows == 0) {
            for (int i = 0; i < have / wrows; ++i)
                for (int k = 0; k < wrows; ++k) weights[i * wrows + k] = w[k];
        } else domainError("weights badly dimensioned");
    }
    void autoSpan() // costFunctions.cpp:36-42,114-120,164-171 (TargetCost: base no-op :20-22)
    {
        if (kind == ORC_COST_TARGET) return;
        int max_dim = std::max(int(weights.size()), int(p.size()));
        if (kind != ORC_COST_CONTROL) max_dim = std::max(max_dim, M.r);
        if (kind != ORC_COST_TRAJECTORY) max_dim = std::max(max_dim, N.r);
        if (kind == ORC_COST_TRAJECTORY) spanMatrix(M, max_dim);
        if (kind == ORC_COST_MIXED) spanMatrix(M, max_dim, 1);
        if (kind == ORC_COST_CONTROL || kind == ORC_COST_MIXED) spanMatrix(N, max_dim);
        spanVector(p, max_dim);
        spanVector(weights, max_dim);
    }
    void initializeCost(const PreviewSystem& ps)
    {
        Q.resize(ps.fullUDim, ps.fullUDim); c.assign(ps.fullUDim, 0.0); // :24-30 (+ zeroing :52-55 etc.)
        E.resize(ps.xDim, ps.fullUDim); f.assign(ps.fullUDim, 0.0);
        switch (kind) {
        case ORC_COST_TRAJECTORY: // :44-61
            if (M.r != int(p.size())) domainError("TrajectoryCost: M/p rows");
            if (M.c == ps.xDim) {} else if (M.c == ps.fullXDim) fullSizeEntry = true;
            else domainError("TrajectoryCost: M cols");
            break;
        case ORC_COST_TARGET: // :88-98
            if (M.r != int(p.size())) domainError("TargetCost: M/p rows");
            if (M.c != ps.xDim) domainError("TargetCost: M cols");
            break;
        case ORC_COST_CONTROL: // :122-137
            if (N.r != int(p.size())) domainError("ControlCost: N/p rows");
            if (N.c == ps.uDim) {} else if (N.c == ps.fullUDim) fullSizeEntry = true;
            else domainError("ControlCost: N cols");
            break;
        case ORC_COST_MIXED: // :173-193
            if (M.r != int(p.size())) domainError("MixedCost: M/p rows");
            if (N.r != int(p.size())) domainError("MixedCost: N/p rows");
            if (M.c == ps.xDim && N.c == ps.uDim) {}
            else if (M.c == ps.fullXDim && N.c == ps.fullUDim) fullSizeEntry = true;
            else domainError("MixedCost: M/N cols");
            break;
        default: domainError("unknown cost kind");
        }
        if (int(weights.size()) != int(p.size())) domainError("weights/p rows");
    }
    // Q += tmp' W tmp ; E += MPhi' W tmp ; f += res' W tmp   (assign==true overwrites)
    void accumulate(const Mat& tmp, const Mat& MPhi, const Vec& res, bool assign)
    {
        int n = tmp.c, r = tmp.r;
        if (assign) { Q.setZero(); E.setZero(); std::fill(f.begin(), f.end(), 0.0); }
        for (int b = 0; b < n; ++b) {
            for (int a = 0; a < n; ++a) {
                double s = 0;
                for (int l = 0; l < r; ++l) s += tmp(l, a) * weights[l] * tmp(l, b);
                Q(a, b) += s;
            }
            for (int a = 0; a < MPhi.c; ++a) {
                double s = 0;
                for (int l = 0; l < r; ++l) s += MPhi(l, a) * weights[l] * tmp(l, b);
                E(a, b) += s;
            }
            double s = 0;
            for (int l = 0; l < r; ++l) s += res[l] * weights[l] * tmp(l, b);
            f[b] += s;
        }
    }
    void finishC(const PreviewSystem& ps, bool plusEq)
    {
        for (int b = 0; b < ps.fullUDim; ++b) {
            double s = 0;
            for (int a = 0; a < ps.xDim; ++a) s += E(a, b) * ps.x0[a];
            if (plusEq) c[b] += s + f[b]; else c[b] = s + f[b];
        }
    }
    void update(const PreviewSystem& ps)
    {
        const int nx = ps.xDim, nu = ps.uDim, n = ps.fullUDim;
        Mat tmp, MPhi;
        Vec res;
        auto residual = [&](int rowOff, int rows, int cols, int xiOff) { // M*xi(seg) - p
            res.assign(rows, 0.0);
            for (int l = 0; l < rows; ++l) {
                double s = 0;
                for (int k = 0; k < cols; ++k) s += M(l, k) * ps.xi[xiOff + k];
                res[l] = s - p[rowOff + l];
            }
        };
        switch (kind) {
        case ORC_COST_TRAJECTORY: // :63-82
            if (fullSizeEntry) {
                mulBlock(M, 0, 0, ps.Psi, 0, 0, M.r, ps.fullXDim, n, tmp);
                mulBlock(M, 0, 0, ps.Phi, 0, 0, M.r, ps.fullXDim, nx, MPhi);
                residual(0, M.r, ps.fullXDim, 0);
                accumulate(tmp, MPhi, res, true);
                finishC(ps, false);
            } else {
                for (int i = 0; i < ps.nrXStep; ++i) { // `+=` : quirk Q2
                    mulBlock(M, 0, 0, ps.Psi, i * nx, 0, M.r, nx, n, tmp);
                    mulBlock(M, 0, 0, ps.Phi, i * nx, 0, M.r, nx, nx, MPhi);
                    residual(0, M.r, nx, i * nx);
                    accumulate(tmp, MPhi, res, false);
                }
                finishC(ps, false);
            }
            break;
        case ORC_COST_TARGET: // :100-108
            mulBlock(M, 0, 0, ps.Psi, ps.fullXDim - nx, 0, M.r, nx, n, tmp);
            mulBlock(M, 0, 0, ps.Phi, ps.fullXDim - nx, 0, M.r, nx, nx, MPhi);
            residual(0, M.r, nx, ps.fullXDim - nx);
            accumulate(tmp, MPhi, res, true);
            finishC(ps, false);
            break;
        case ORC_COST_CONTROL: // :139-158
            if (fullSizeEntry) {
                MPhi.resize(N.r, nx);
                res.assign(N.r, 0.0);
                for (int l = 0; l < N.r; ++l) res[l] = -p[l];
                accumulate(N, MPhi, res, true);
                c = f;
            } else {
                Mat mat(nu, nu);
                Vec vec(nu, 0.0);
                for (int b = 0; b < nu; ++b) {
                    for (int a = 0; a < nu; ++a) {
                        double s = 0;
                        for (int l = 0; l < N.r; ++l) s += N(l, a) * weights[l] * N(l, b);
                        mat(a, b) = s;
                    }
                    double s = 0;
                    for (int l = 0; l < N.r; ++l) s += -p[l] * weights[l] * N(l, b);
                    vec[b] = s;
                }
                for (int i = 0; i < ps.nrUStep; ++i) {
                    for (int b = 0; b < nu; ++b) {
                        for (int a = 0; a < nu; ++a) Q(i * nu + a, i * nu + b) = mat(a, b);
                        for (int a = 0; a < nx; ++a) E(a, i * nu + b) = 0.0;
                        f[i * nu + b] = vec[b];
                        c[i * nu + b] = f[i * nu + b];
                    }
                }
            }
            break;
        case ORC_COST_MIXED: // :195-215
            if (fullSizeEntry) {
                mulBlock(M, 0, 0, ps.Psi, 0, 0, M.r, ps.fullXDim, n, tmp);
                for (size_t k = 0; k < tmp.a.size(); ++k) tmp.a[k] += N.a[k];
                mulBlock(M, 0, 0, ps.Phi, 0, 0, M.r, ps.fullXDim, nx, MPhi);
                residual(0, M.r, ps.fullXDim, 0);
                accumulate(tmp, MPhi, res, true);
                finishC(ps, false);
            } else {
                for (int i = 0; i < ps.nrUStep; ++i) { // x_N not penalised: quirk Q6
                    mulBlock(M, 0, 0, ps.Psi, i * nx, 0, M.r, nx, n, tmp);
                    for (int cc = 0; cc < nu; ++cc)
                        for (int l = 0; l < M.r; ++l) tmp(l, i * nu + cc) += N(l, cc);
                    mulBlock(M, 0, 0, ps.Phi, i * nx, 0, M.r, nx, nx, MPhi);
                    residual(0, M.r, nx, i * nx);
                    accumulate(tmp, MPhi, res, false);
                }
                finishC(ps, true); // `c_ +=` (:213)
            }
            break;
        }
    }
};

/* ---------------------------------------------------------------------------------------------
 * Constraints -- reference src/constraints.cpp:16-372, include/constraints.h:242-255
 * ------------------------------------------------------------------------------------------- */
struct Constraint {
    int kind;
    bool isIneq = true, fullSizeEntry = false, hasBeenInitialized = false;
    int nrConstr = 0;
    Mat E, G, A, Y;
    Vec f, b, z, lower, upper, lb, ub;
    std::vector<int> lowerLines, upperLines;

    explicit Constraint(const orc_constraint& s) : kind(s.kind), isIneq(s.is_ineq != 0)
    {
        switch (kind) {
        case ORC_CSTR_TRAJECTORY: E = Mat(s.rows, s.colsE, s.E); f.assign(s.f, s.f + s.rows); break;
        case ORC_CSTR_CONTROL: G = Mat(s.rows, s.colsG, s.G); f.assign(s.f, s.f + s.rows); break;
        case ORC_CSTR_MIXED:
            E = Mat(s.rows, s.colsE, s.E); G = Mat(s.rows, s.colsG, s.G); f.assign(s.f, s.f + s.rows);
            break;
        case ORC_CSTR_TRAJECTORY_BOUND:
            isIneq = true;
            lower.assign(s.lower, s.lower + s.rows); upper.assign(s.upper, s.upper + s.rows);
            selectLines(true, true); // ctor, constraints.h:247-254
            break;
        case ORC_CSTR_CONTROL_BOUND:
            lower.assign(s.lower, s.lower + s.rows); upper.assign(s.upper, s.upper + s.rows);
            break;
        default: domainError("unknown constraint kind");
        }
        if (s.autospan) autoSpan();
    }
    void selectLines(bool lo, bool up)
    {
        if (lo) {
            lowerLines.clear();
            for (int l = 0; l < int(lower.size()); ++l) if (lower[l] != -kInf) lowerLines.push_back(l);
        }
        if (up) {
            upperLines.clear();
            for (int l = 0; l < int(upper.size()); ++l) if (upper[l] != kInf) upperLines.push_back(l);
        }
    }
    void autoSpan()
    {
        switch (kind) {
        case ORC_CSTR_TRAJECTORY: { // :38-43
            int md = std::max(E.r, int(f.size())); spanMatrix(E, md); spanVector(f, md);
        } break;
        case ORC_CSTR_CONTROL: { // :99-104
            int md = std::max(G.r, int(f.size())); spanMatrix(G, md); spanVector(f, md);
        } break;
        case ORC_CSTR_MIXED: { // :163-169
            int md = std::max(int(f.size()), std::max(E.r, G.r));
            spanMatrix(E, md, 1); spanMatrix(G, md); spanVector(f, md);
        } break;
        case ORC_CSTR_TRAJECTORY_BOUND: { // :241-261 (the rows() tests there are evaluated after spanning)
            int md = int(std::max(lower.size(), upper.size()));
            spanVector(lower, md); spanVector(upper, md);
            if (int(lower.size()) != md) selectLines(true, false);
            if (int(upper.size()) != md) selectLines(false, true);
        } break;
        case ORC_CSTR_CONTROL_BOUND: { // :326-331
            int md = int(std::max(lower.size(), upper.size()));
            spanVector(lower, md); spanVector(upper, md);
        } break;
        }
    }
    void initializeConstraint(const PreviewSystem& ps)
    {
        switch (kind) {
        case ORC_CSTR_TRAJECTORY: // :45-64
            if (E.r != int(f.size())) domainError("TrajectoryConstraint: E/f rows");
            if (E.c == ps.xDim) nrConstr = E.r * ps.nrXStep;
            else if (E.c == ps.fullXDim) { fullSizeEntry = true; nrConstr = E.r; }
            else domainError("TrajectoryConstraint: E cols");
            break;
        case ORC_CSTR_CONTROL: // :106-135
            if (hasBeenInitialized) runtimeError("ControlConstraint initialized twice");
            if (G.r != int(f.size())) domainError("ControlConstraint: G/f rows");
            if (G.c == ps.uDim) nrConstr = G.r * ps.nrUStep;
            else if (G.c == ps.fullUDim) { fullSizeEntry = true; nrConstr = G.r; }
            else domainError("ControlConstraint: G cols");
            break;
        case ORC_CSTR_MIXED: // :171-195
            if (E.r != int(f.size())) domainError("MixedConstraint: E/f rows");
            if (G.r != int(f.size())) domainError("MixedConstraint: G/f rows");
            if (E.c == ps.xDim && G.c == ps.uDim) nrConstr = E.r * ps.nrUStep;
            else if (E.c == ps.fullXDim && G.c == ps.fullUDim) { fullSizeEntry = true; nrConstr = E.r; }
            else domainError("MixedConstraint: E/G cols");
            break;
        case ORC_CSTR_TRAJECTORY_BOUND: // :263-282
            if (lower.size() != upper.size()) domainError("TrajectoryBoundConstraint: lower/upper rows");
            if (int(lower.size()) == ps.xDim) nrConstr = int(lowerLines.size() + upperLines.size()) * ps.nrXStep;
            else if (int(lower.size()) == ps.fullXDim) { nrConstr = int(lowerLines.size() + upperLines.size()); fullSizeEntry = true; }
            else domainError("TrajectoryBoundConstraint: rows");
            break;
        case ORC_CSTR_CONTROL_BOUND: // :333-357
            if (hasBeenInitialized) runtimeError("ControlBoundConstraint initialized twice");
            if (lower.size() != upper.size()) domainError("ControlBoundConstraint: lower/upper rows");
            if (int(lower.size()) == ps.uDim) { nrConstr = ps.fullUDim; lb.assign(nrConstr, 0.0); ub.assign(nrConstr, 0.0); }
            else if (int(lower.size()) == ps.fullUDim) { fullSizeEntry = true; nrConstr = int(lower.size()); lb = lower; ub = upper; }
            else domainError("ControlBoundConstraint: rows");
            break;
        }
        if (kind != ORC_CSTR_CONTROL_BOUND) {
            A.resize(nrConstr, ps.fullUDim); Y.resize(nrConstr, ps.xDim);
            b.assign(nrConstr, 0.0); z.assign(nrConstr, 0.0);
            if (kind == ORC_CSTR_CONTROL && fullSizeEntry) { A = G; b = f; z = b; } // :122-126,131
        }
        hasBeenInitialized = true;
    }
    // rows [row0,row0+nl): A = E.block(er,ec,nl,k)*Psi.block(xr,..) ; Y ; z = f - E xi ; b = z - Y x0
    void stateRows(const PreviewSystem& ps, int row0, int nl, int er, int ec, int k, int xr, bool doA)
    {
        const int n = ps.fullUDim, nx = ps.xDim;
        if (doA)
            for (int j = 0; j < n; ++j)
                for (int l = 0; l < nl; ++l) {
                    double s = 0;
                    for (int q = 0; q < k; ++q) s += E(er + l, ec + q) * ps.Psi(xr + q, j);
                    A(row0 + l, j) = s;
                }
        for (int j = 0; j < nx; ++j)
            for (int l = 0; l < nl; ++l) {
                double s = 0;
                for (int q = 0; q < k; ++q) s += E(er + l, ec + q) * ps.Phi(xr + q, j);
                Y(row0 + l, j) = s;
            }
        for (int l = 0; l < nl; ++l) {
            double s = 0;
            for (int q = 0; q < k; ++q) s += E(er + l, ec + q) * ps.xi[xr + q];
            z[row0 + l] = f[er + l] - s;
        }
        finishB(ps, row0, nl);
    }
    void finishB(const PreviewSystem& ps, int row0, int nl)
    {
        for (int l = 0; l < nl; ++l) {
            double s = 0;
            for (int q = 0; q < ps.xDim; ++q) s += Y(row0 + l, q) * ps.x0[q];
            b[row0 + l] = z[row0 + l] - s;
        }
    }
    void update(const PreviewSystem& ps)
    {
        const int nx = ps.xDim, nu = ps.uDim;
        switch (kind) {
        case ORC_CSTR_TRAJECTORY: // :66-84
            if (fullSizeEntry) stateRows(ps, 0, E.r, 0, 0, ps.fullXDim, 0, true);
            else
                for (int i = 0; i < ps.nrXStep; ++i) stateRows(ps, i * E.r, E.r, 0, 0, nx, i * nx, true);
            break;
        case ORC_CSTR_CONTROL: // :137-148
            if (!fullSizeEntry) {
                int nl = G.r;
                for (int i = 0; i < ps.nrUStep; ++i) {
                    for (int cc = 0; cc < nu; ++cc)
                        for (int l = 0; l < nl; ++l) A(i * nl + l, i * nu + cc) = G(l, cc);
                    for (int l = 0; l < nl; ++l) b[i * nl + l] = f[l];
                }
                Y.setZero();
                z = b;
            }
            break;
        case ORC_CSTR_MIXED: // :197-226
            if (fullSizeEntry) {
                stateRows(ps, 0, E.r, 0, 0, ps.fullXDim, 0, true);
                for (size_t k = 0; k < A.a.size(); ++k) A.a[k] += G.a[k];
            } else {
                int nl = E.r;
                for (int cc = 0; cc < nu; ++cc)
                    for (int l = 0; l < nl; ++l) A(l, cc) = G(l, cc);
                for (int cc = 0; cc < nx; ++cc)
                    for (int l = 0; l < nl; ++l) Y(l, cc) = E(l, cc);
                for (int l = 0; l < nl; ++l) z[l] = f[l];
                finishB(ps, 0, nl);
                for (int i = 1; i < ps.nrUStep; ++i) {
                    // first block column by product, the rest by Toeplitz shift copy (:214-219)
                    for (int cc = 0; cc < nu; ++cc)
                        for (int l = 0; l < nl; ++l) {
                            double s = 0;
                            for (int q = 0; q < nx; ++q) s += E(l, q) * ps.Psi(i * nx + q, cc);
                            A(i * nl + l, cc) = s;
                        }
                    for (int j = 1; j <= i; ++j)
                        for (int cc = 0; cc < nu; ++cc)
                            for (int l = 0; l < nl; ++l)
                                A(i * nl + l, j * nu + cc) = A((i - 1) * nl + l, (j - 1) * nu + cc);
                    stateRows(ps, i * nl, nl, 0, 0, nx, i * nx, false);
                }
            }
            break;
        case ORC_CSTR_TRAJECTORY_BOUND: { // :284-315 -- NB lower rows are NOT negated (quirk Q1)
            int row = 0;
            for (int pass = 0; pass < 2; ++pass) {
                const std::vector<int>& lines = pass == 0 ? lowerLines : upperLines;
                const Vec& bound = pass == 0 ? lower : upper;
                for (int step = 0; step < ps.nrXStep; ++step) {
                    for (int line : lines) {
                        int src = line + nx * step;
                        for (int j = 0; j < ps.fullUDim; ++j) A(row, j) = ps.Psi(src, j);
                        for (int j = 0; j < nx; ++j) Y(row, j) = ps.Phi(src, j);
                        z[row] = bound[line] - ps.xi[src];
                        finishB(ps, row, 1);
                        ++row;
                    }
                    if (fullSizeEntry) break;
                }
            }
        } break;
        case ORC_CSTR_CONTROL_BOUND: // :359-367
            if (!fullSizeEntry)
                for (int i = 0; i < ps.nrUStep; ++i)
                    for (int k = 0; k < nu; ++k) { ub[i * nu + k] = upper[k]; lb[i * nu + k] = lower[k]; }
            break;
        }
    }
};

/* ---------------------------------------------------------------------------------------------
 * LINPACK dpofa / dposl / dpori as used by qpgen2 [external: eigen-quadprog, unpinned; SURVEY 3.3]
 * a is column-major n x n with leading dimension lda; upper triangle holds the factor.
 * ------------------------------------------------------------------------------------------- */
int dpofa(double* a, int lda, int n)
{
    for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int k = 0; k < j; ++k) {
            double t = a[size_t(j) * lda + k];
            for (int i = 0; i < k; ++i) t -= a[size_t(k) * lda + i] * a[size_t(j) * lda + i];
            t /= a[size_t(k) * lda + k];
            a[size_t(j) * lda + k] = t;
            s += t * t;
        }
        s = a[size_t(j) * lda + j] - s;
        if (s <= 0.0) return j + 1;
        a[size_t(j) * lda + j] = std::sqrt(s);
    }
    return 0;
}
void dposl(const double* a, int lda, int n, double* b)
{
    for (int k = 0; k < n; ++k) { // R' y = b
        double t = 0;
        for (int i = 0; i < k; ++i) t += a[size_t(k) * lda + i] * b[i];
        b[k] = (b[k] - t) / a[size_t(k) * lda + k];
    }
    for (int kb = 0; kb < n; ++kb) { // R x = y
        int k = n - 1 - kb;
        b[k] /= a[size_t(k) * lda + k];
        double t = -b[k];
        for (int i = 0; i < k; ++i) b[i] += t * a[size_t(k) * lda + i];
    }
}
void dpori(double* a, int lda, int n)
{
    for (int k = 0; k < n; ++k) { // inverse of upper-triangular R, in place
        a[size_t(k) * lda + k] = 1.0 / a[size_t(k) * lda + k];
        double t = -a[size_t(k) * lda + k];
        for (int i = 0; i < k; ++i) a[size_t(k) * lda + i] *= t;
        for (int j = k + 1; j < n; ++j) {
            t = a[size_t(j) * lda + k];
            a[size_t(j) * lda + k] = 0.0;
            for (int i = 0; i <= k; ++i) a[size_t(j) * lda + i] += t * a[size_t(k) * lda + i];
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * qpgen2 -- Goldfarb-Idnani dual active set, as in quadprog's Fortran [external, SURVEY 3.3].
 *   min 1/2 x'Dx - d'x   s.t.  A' x >= b   (first meq are equalities)
 * dmat (n x n, destroyed: holds J = R^-1 on exit), dvec (n, destroyed), amat (n x q, one COLUMN per
 * constraint; equality columns may be sign-flipped in place), bvec (q).  All index arithmetic is
 * kept 1-based inside `work` like the Fortran so that the packed R layout is identical.
 * ------------------------------------------------------------------------------------------- */
int qpgen2(double* dmat, double* dvec, int n, double* sol, double* lagr, double* crval_out,
    double* amat, double* bvec, int q, int meq, int* iact, int* nact_out, int* iter)
{
    const int fdd = n, fda = n;
    auto D = [&](int i, int j) -> double& { return dmat[size_t(j - 1) * fdd + (i - 1)]; };
    auto Am = [&](int i, int j) -> double& { return amat[size_t(j - 1) * fda + (i - 1)]; };
    const int r = std::min(n, q);
    int l = 2 * n + (r * (r + 5)) / 2 + 2 * q + 1;
    std::vector<double> workv(size_t(l) + 2, 0.0);
    double* work = workv.data(); // 1-based: work[1..l]

    double vsmall = 1.0e-60;
    for (;;) {
        vsmall += vsmall;
        volatile double tmpa = 1.0 + 0.1 * vsmall;
        volatile double tmpb = 1.0 + 0.2 * vsmall;
        if (tmpa <= 1.0) continue;
        if (tmpb <= 1.0) continue;
        break;
    }
    for (int i = 1; i <= n; ++i) work[i] = dvec[i - 1];
    for (int i = 1; i <= q; ++i) { iact[i - 1] = 0; if (lagr) lagr[i - 1] = 0.0; }

    int info = dpofa(dmat, fdd, n);
    if (info != 0) { *nact_out = 0; iter[0] = iter[1] = 0; *crval_out = 0; return 2; }
    dposl(dmat, fdd, n, dvec);
    dpori(dmat, fdd, n);

    double crval = 0.0;
    for (int j = 1; j <= n; ++j) {
        sol[j - 1] = dvec[j - 1];
        crval += work[j] * sol[j - 1];
        work[j] = 0.0;
        for (int i = j + 1; i <= n; ++i) D(i, j) = 0.0;
    }
    crval = -crval / 2.0;
    int ierr = 0;

    const int iwzv = n, iwrv = iwzv + n, iwuv = iwrv + r, iwrm = iwuv + r + 1;
    const int iwsv = iwrm + (r * (r + 1)) / 2, iwnbv = iwsv + q;

    for (int i = 1; i <= q; ++i) {
        double sum = 0.0;
        for (int j = 1; j <= n; ++j) sum += Am(j, i) * Am(j, i);
        work[iwnbv + i] = std::sqrt(sum);
    }
    int nact = 0;
    iter[0] = 0; iter[1] = 0;
    int nvl = 0, it1 = 0;
    double t1 = 0, tt, sum, temp, gc, gs, nu;
    bool t1inf, t2min;

L50: // start a new iteration
    iter[0] += 1;
    l = iwsv;
    for (int i = 1; i <= q; ++i) {
        l += 1;
        sum = -bvec[i - 1];
        for (int j = 1; j <= n; ++j) sum += Am(j, i) * sol[j - 1];
        if (std::fabs(sum) < vsmall) sum = 0.0;
        if (i > meq) work[l] = sum;
        else {
            work[l] = -std::fabs(sum);
            if (sum > 0.0) {
                for (int j = 1; j <= n; ++j) Am(j, i) = -Am(j, i);
                bvec[i - 1] = -bvec[i - 1];
            }
        }
    }
    for (int i = 1; i <= nact; ++i) work[iwsv + iact[i - 1]] = 0.0;
    nvl = 0;
    temp = 0.0;
    for (int i = 1; i <= q; ++i) {
        if (work[iwsv + i] < temp * work[iwnbv + i]) {
            nvl = i;
            temp = work[iwsv + i] / work[iwnbv + i];
        }
    }
    if (nvl == 0) {
        if (lagr) for (int i = 1; i <= nact; ++i) lagr[iact[i - 1] - 1] = work[iwuv + i];
        goto L999;
    }

L55: // d = J' n+
    for (int i = 1; i <= n; ++i) {
        sum = 0.0;
        for (int j = 1; j <= n; ++j) sum += D(j, i) * Am(j, nvl);
        work[i] = sum;
    }
    // z = J_2 d_2
    for (int i = 1; i <= n; ++i) work[iwzv + i] = 0.0;
    for (int j = nact + 1; j <= n; ++j)
        for (int i = 1; i <= n; ++i) work[iwzv + i] += D(i, j) * work[j];
    // r = R^-1 d_1
    t1inf = true;
    for (int i = nact; i >= 1; --i) {
        sum = work[i];
        l = iwrm + (i * (i + 3)) / 2;
        int l1 = l - i;
        for (int j = i + 1; j <= nact; ++j) {
            sum -= work[l] * work[iwrv + j];
            l += j;
        }
        sum /= work[l1];
        work[iwrv + i] = sum;
        if (iact[i - 1] <= meq) continue;
        if (sum <= 0.0) continue;
        t1inf = false;
        it1 = i;
    }
    if (!t1inf) {
        t1 = work[iwuv + it1] / work[iwrv + it1];
        for (int i = 1; i <= nact; ++i) {
            if (iact[i - 1] <= meq) continue;
            if (work[iwrv + i] <= 0.0) continue;
            temp = work[iwuv + i] / work[iwrv + i];
            if (temp < t1) { t1 = temp; it1 = i; }
        }
    }
    sum = 0.0;
    for (int i = iwzv + 1; i <= iwzv + n; ++i) sum += work[i] * work[i];
    if (std::fabs(sum) <= vsmall) {
        if (t1inf) { ierr = 1; goto L999; }
        for (int i = 1; i <= nact; ++i) work[iwuv + i] -= t1 * work[iwrv + i];
        work[iwuv + nact + 1] += t1;
        goto L700;
    } else {
        sum = 0.0;
        for (int i = 1; i <= n; ++i) sum += work[iwzv + i] * Am(i, nvl);
        tt = -work[iwsv + nvl] / sum;
        t2min = true;
        if (!t1inf) {
            if (t1 < tt) { tt = t1; t2min = false; }
        }
        for (int i = 1; i <= n; ++i) sol[i - 1] += tt * work[iwzv + i];
        crval += tt * sum * (tt / 2.0 + work[iwuv + nact + 1]);
        for (int i = 1; i <= nact; ++i) work[iwuv + i] -= tt * work[iwrv + i];
        work[iwuv + nact + 1] += tt;
        if (t2min) {
            nact += 1;
            iact[nact - 1] = nvl;
            l = iwrm + ((nact - 1) * nact) / 2 + 1;
            for (int i = 1; i <= nact - 1; ++i) { work[l] = work[i]; l += 1; }
            if (nact == n) {
                work[l] = work[n];
            } else {
                for (int i = n; i >= nact + 1; --i) {
                    if (work[i] == 0.0) continue;
                    gc = std::max(std::fabs(work[i - 1]), std::fabs(work[i]));
                    gs = std::min(std::fabs(work[i - 1]), std::fabs(work[i]));
                    temp = std::copysign(gc * std::sqrt(1 + gs * gs / (gc * gc)), work[i - 1]);
                    if (work[i - 1] == 0.0) temp = std::fabs(temp); // Fortran SIGN(a, +-0) -> +|a|
                    gc = work[i - 1] / temp;
                    gs = work[i] / temp;
                    if (gc == 1.0) continue;
                    if (gc == 0.0) {
                        work[i - 1] = gs * temp;
                        for (int j = 1; j <= n; ++j) { temp = D(j, i - 1); D(j, i - 1) = D(j, i); D(j, i) = temp; }
                    } else {
                        work[i - 1] = temp;
                        nu = gs / (1.0 + gc);
                        for (int j = 1; j <= n; ++j) {
                            temp = gc * D(j, i - 1) + gs * D(j, i);
                            D(j, i) = nu * (D(j, i - 1) + temp) - D(j, i);
                            D(j, i - 1) = temp;
                        }
                    }
                }
                work[l] = work[nact];
            }
        } else {
            sum = -bvec[nvl - 1];
            for (int j = 1; j <= n; ++j) sum += sol[j - 1] * Am(j, nvl);
            if (nvl > meq) work[iwsv + nvl] = sum;
            else {
                work[iwsv + nvl] = -std::fabs(sum);
                if (sum > 0.0) {
                    for (int j = 1; j <= n; ++j) Am(j, nvl) = -Am(j, nvl);
                    bvec[nvl - 1] = -bvec[nvl - 1];
                }
            }
            goto L700;
        }
    }
    goto L50;

L700: // drop constraint it1
    if (it1 == nact) goto L799;
L797: {
    l = iwrm + (it1 * (it1 + 1)) / 2 + 1;
    int l1 = l + it1;
    if (work[l1] == 0.0) goto L798;
    gc = std::max(std::fabs(work[l1 - 1]), std::fabs(work[l1]));
    gs = std::min(std::fabs(work[l1 - 1]), std::fabs(work[l1]));
    temp = std::copysign(gc * std::sqrt(1 + gs * gs / (gc * gc)), work[l1 - 1]);
    if (work[l1 - 1] == 0.0) temp = std::fabs(temp);
    gc = work[l1 - 1] / temp;
    gs = work[l1] / temp;
    if (gc == 1.0) goto L798;
    if (gc == 0.0) {
        for (int i = it1 + 1; i <= nact; ++i) {
            temp = work[l1 - 1]; work[l1 - 1] = work[l1]; work[l1] = temp;
            l1 += i;
        }
        for (int i = 1; i <= n; ++i) { temp = D(i, it1); D(i, it1) = D(i, it1 + 1); D(i, it1 + 1) = temp; }
    } else {
        nu = gs / (1.0 + gc);
        for (int i = it1 + 1; i <= nact; ++i) {
            temp = gc * work[l1 - 1] + gs * work[l1];
            work[l1] = nu * (work[l1 - 1] + temp) - work[l1];
            work[l1 - 1] = temp;
            l1 += i;
        }
        for (int i = 1; i <= n; ++i) {
            temp = gc * D(i, it1) + gs * D(i, it1 + 1);
            D(i, it1 + 1) = nu * (D(i, it1) + temp) - D(i, it1 + 1);
            D(i, it1) = temp;
        }
    }
}
L798: {
    int l1 = l - it1;
    for (int i = 1; i <= it1; ++i) { work[l1] = work[l]; l += 1; l1 += 1; }
    work[iwuv + it1] = work[iwuv + it1 + 1];
    iact[it1 - 1] = iact[it1];
    it1 += 1;
    if (it1 < nact) goto L797;
}
L799:
    work[iwuv + nact] = work[iwuv + nact + 1];
    work[iwuv + nact + 1] = 0.0;
    iact[nact - 1] = 0;
    nact -= 1;
    iter[1] += 1;
    goto L55;

L999:
    *nact_out = nact;
    *crval_out = crval;
    return ierr;
}

/* ---------------------------------------------------------------------------------------------
 * QuadProgDenseSolver::SI_solve (src/QuadProgSolver.cpp:45-72) + Eigen::QuadProgDense::solve
 * [external]: densify bounds to [Aineq; I; -I] x <= [bineq; XU; -XL], then map onto qpgen2:
 *   D=Q, d=-c, A=[Aeq' , -ineqMat'], b0=[beq; -ineqVec].
 * ------------------------------------------------------------------------------------------- */
int quadprogSolve(int n, int meq, int m, const double* Q, const double* c, const double* Aeq,
    const double* beq, const double* Aineq, const double* bineq, const double* XL, const double* XU,
    double* x, int* iact, int* nact, int* iter, double* lagr, double* crval)
{
    const int nin = m + 2 * n, q = meq + nin;
    std::vector<double> ineqMat(size_t(nin) * n, 0.0), ineqVec(nin);
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < m; ++i) ineqMat[size_t(j) * nin + i] = Aineq[size_t(j) * m + i];
        ineqMat[size_t(j) * nin + m + j] = 1.0;
        ineqMat[size_t(j) * nin + m + n + j] = -1.0;
    }
    for (int i = 0; i < m; ++i) ineqVec[i] = bineq[i];
    for (int i = 0; i < n; ++i) { ineqVec[m + i] = XU[i]; ineqVec[m + n + i] = -XL[i]; }

    std::vector<double> D(Q, Q + size_t(n) * n), dv(n), A(size_t(n) * q), b0(q);
    for (int i = 0; i < n; ++i) dv[i] = -c[i];
    for (int k = 0; k < meq; ++k) {
        for (int j = 0; j < n; ++j) A[size_t(k) * n + j] = Aeq[size_t(j) * meq + k];
        b0[k] = beq[k];
    }
    for (int k = 0; k < nin; ++k) {
        for (int j = 0; j < n; ++j) A[size_t(meq + k) * n + j] = -ineqMat[size_t(j) * nin + k];
        b0[meq + k] = -ineqVec[k];
    }
    std::vector<int> iactv(q > 0 ? q : 1, 0);
    int it[2] = { 0, 0 }, na = 0;
    double cr = 0;
    int ierr = qpgen2(D.data(), dv.data(), n, x, lagr, &cr, A.data(), b0.data(), q, meq, iactv.data(), &na, it);
    if (iact) std::copy(iactv.begin(), iactv.begin() + q, iact);
    if (nact) *nact = na;
    if (iter) { iter[0] = it[0]; iter[1] = it[1]; }
    if (crval) *crval = cr;
    return ierr;
}

/* Eigen's MatrixXd::inverse() for dynamic sizes = PartialPivLU (quirk Q7): restated as LU with
 * partial pivoting + solve against the identity. */
Mat luInverse(const Mat& Min)
{
    int n = Min.r;
    Mat LU = Min, inv(n, n);
    std::vector<int> perm(n);
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double best = std::fabs(LU(k, k));
        for (int i = k + 1; i < n; ++i) if (std::fabs(LU(i, k)) > best) { best = std::fabs(LU(i, k)); piv = i; }
        if (piv != k) {
            for (int j = 0; j < n; ++j) std::swap(LU(k, j), LU(piv, j));
            std::swap(perm[k], perm[piv]);
        }
        for (int i = k + 1; i < n; ++i) LU(i, k) /= LU(k, k);
        for (int j = k + 1; j < n; ++j) {
            double u = LU(k, j);
            for (int i = k + 1; i < n; ++i) LU(i, j) -= LU(i, k) * u;
        }
    }
    for (int col = 0; col < n; ++col) {
        Vec y(n);
        for (int i = 0; i < n; ++i) y[i] = perm[i] == col ? 1.0 : 0.0;
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < i; ++k) y[i] -= LU(i, k) * y[k];
        for (int i = n - 1; i >= 0; --i) {
            for (int k = i + 1; k < n; ++k) y[i] -= LU(i, k) * y[k];
            y[i] /= LU(i, i);
        }
        for (int i = 0; i < n; ++i) inv(i, col) = y[i];
    }
    return inv;
}

/* ---------------------------------------------------------------------------------------------
 * LMPC / InitialStateLMPC -- reference src/LMPC.cpp:79-101,199-286; src/InitialStateLMPC.cpp:52-128
 * ------------------------------------------------------------------------------------------- */
struct Controller {
    PreviewSystem ps;
    std::vector<std::unique_ptr<Cost>> costs;
    std::vector<std::unique_ptr<Constraint>> cstrs;
    std::vector<Constraint*> eq, ineq, bound;
    bool initialState = false;
    int nvar = 0, nrEq = 0, nrIneq = 0;
    Mat Q, Aeq, Aineq, R;
    Vec c, beq, bineq, lb, ub, r, x0lb, x0ub, result, control, trajectory;
    std::vector<int> iact;
    int nact = 0, iter[2] = { 0, 0 }, fail = 0;
    Vec lagr;
    double crval = 0;

    void setup(const orc_problem& p)
    {
        ps.system(p.nx, p.nu, p.N, p.A, p.B, p.d, p.x0);
        initialState = p.initial_state != 0;
        nvar = initialState ? ps.xDim + ps.fullUDim : ps.fullUDim;
        if (initialState) { // InitialStateLMPC.cpp:19-28
            R.resize(ps.xDim, ps.xDim);
            if (p.R) R = Mat(ps.xDim, ps.xDim, p.R);
            r.assign(ps.xDim, 0.0);
            if (p.r) r.assign(p.r, p.r + ps.xDim);
            x0lb = ps.x0; x0ub = ps.x0;
            if (p.x0lb) x0lb.assign(p.x0lb, p.x0lb + ps.xDim);
            if (p.x0ub) x0ub.assign(p.x0ub, p.x0ub + ps.xDim);
        }
        for (int i = 0; i < p.ncost; ++i) { // LMPC::addCost (LMPC.cpp:118-122)
            costs.emplace_back(new Cost(p.costs[i]));
            costs.back()->initializeCost(ps);
        }
        for (int i = 0; i < p.ncstr; ++i) { // LMPC::addConstraint (:124-128,173-197)
            cstrs.emplace_back(new Constraint(p.cstrs[i]));
            Constraint* cs = cstrs.back().get();
            cs->initializeConstraint(ps);
            if (cs->kind == ORC_CSTR_CONTROL_BOUND) bound.push_back(cs);
            else if (cs->isIneq) ineq.push_back(cs);
            else eq.push_back(cs);
        }
    }
    void build()
    {
        // LMPC::updateSystem (:225-248)
        Q.resize(nvar, nvar);
        for (int i = 0; i < nvar; ++i) Q(i, i) = 1e-6;
        c.assign(nvar, 0.0);
        ps.updateSystem();
        nrEq = 0; nrIneq = 0;
        for (auto* e : eq) nrEq += e->nrConstr;
        for (auto* e : ineq) nrIneq += e->nrConstr;
        Aeq.resize(nrEq, nvar); beq.assign(nrEq, 0.0);
        Aineq.resize(nrIneq, nvar); bineq.assign(nrIneq, 0.0);
        lb.assign(nvar, -std::numeric_limits<double>::max()); // :207-208 (quirk Q4)
        ub.assign(nvar, std::numeric_limits<double>::max());
        for (auto& cs : cstrs) cs->update(ps);
        for (auto& co : costs) co->update(ps);
        const int off = initialState ? ps.xDim : 0, n = ps.fullUDim;
        // makeQPForm (:250-280 / InitialStateLMPC.cpp:77-122)
        for (auto& co : costs) {
            for (int j = 0; j < n; ++j) {
                for (int i = 0; i < n; ++i) Q(off + i, off + j) += co->Q(i, j);
                if (initialState) {
                    for (int i = 0; i < ps.xDim; ++i) Q(i, off + j) += co->E(i, j);
                    c[off + j] += co->f[j];
                } else c[j] += co->c[j];
            }
        }
        auto stack = [&](std::vector<Constraint*>& list, Mat& Aout, Vec& bout) {
            int row = 0;
            for (auto* cs : list) {
                for (int l = 0; l < cs->nrConstr; ++l) {
                    if (initialState) for (int j = 0; j < ps.xDim; ++j) Aout(row + l, j) = cs->Y(l, j);
                    for (int j = 0; j < n; ++j) Aout(row + l, off + j) = cs->A(l, j);
                    bout[row + l] = initialState ? cs->z[l] : cs->b[l];
                }
                row += cs->nrConstr;
            }
        };
        stack(eq, Aeq, beq);
        stack(ineq, Aineq, bineq);
        int row = off;
        for (auto* cs : bound) { // consecutive segments, quirk Q8
            if (row + cs->nrConstr > nvar) runtimeError("bound constraints overflow lb/ub (quirk Q8)");
            for (int l = 0; l < cs->nrConstr; ++l) { lb[row + l] = cs->lb[l]; ub[row + l] = cs->ub[l]; }
            row += cs->nrConstr;
        }
        if (initialState) { // InitialStateLMPC.cpp:112-121
            const int nx = ps.xDim;
            Mat Qb(n, n), Eb(nx, n);
            for (int j = 0; j < n; ++j) {
                for (int i = 0; i < n; ++i) Qb(i, j) = Q(off + i, off + j);
                for (int i = 0; i < nx; ++i) { Eb(i, j) = Q(i, off + j); Q(off + j, i) = Eb(i, j); }
            }
            Mat Qinv = luInverse(Qb), EQ(nx, n);
            for (int j = 0; j < n; ++j)
                for (int k = 0; k < n; ++k) {
                    double v = Qinv(k, j);
                    for (int i = 0; i < nx; ++i) EQ(i, j) += Eb(i, k) * v;
                }
            for (int j = 0; j < nx; ++j)
                for (int i = 0; i < nx; ++i) {
                    double s = 0;
                    for (int k = 0; k < n; ++k) s += EQ(i, k) * Eb(j, k);
                    Q(i, j) = R(i, j) + s;
                }
            for (int i = 0; i < nx; ++i) { c[i] = r[i]; lb[i] = x0lb[i]; ub[i] = x0ub[i]; }
        }
    }
    bool solve()
    {
        const int q = nrEq + nrIneq + 2 * nvar;
        result.assign(nvar, 0.0); iact.assign(q, 0); lagr.assign(q, 0.0);
        fail = quadprogSolve(nvar, nrEq, nrIneq, Q.a.data(), c.data(), Aeq.a.data(), beq.data(), Aineq.a.data(),
            bineq.data(), lb.data(), ub.data(), result.data(), iact.data(), &nact, iter, lagr.data(), &crval);
        if (fail != 0) return false;
        // updateResults (LMPC.cpp:282-286 / InitialStateLMPC.cpp:124-128)
        const int off = initialState ? ps.xDim : 0;
        control.assign(result.begin() + off, result.end());
        const double* xinit = initialState ? result.data() : ps.x0.data();
        trajectory.assign(ps.fullXDim, 0.0);
        for (int i = 0; i < ps.fullXDim; ++i) {
            double s = 0;
            for (int k = 0; k < ps.xDim; ++k) s += ps.Phi(i, k) * xinit[k];
            double s2 = 0;
            for (int k = 0; k < ps.fullUDim; ++k) s2 += ps.Psi(i, k) * control[k];
            trajectory[i] = s + s2 + ps.xi[i];
        }
        return true;
    }
};

template <class T> void copyOut(T* dst, const std::vector<T>& v) { if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(T)); }

int guarded(const std::function<void()>& fn)
{
    try { fn(); return 0; }
    catch (const std::domain_error& e) { g_err = e.what(); return -1; }
    catch (const std::runtime_error& e) { g_err = e.what(); return -2; }
    catch (const std::exception& e) { g_err = e.what(); return -3; }
}

} // namespace

extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }
int orc_hw_threads(void) { unsigned t = std::thread::hardware_concurrency(); return t ? int(t) : 1; }

int orc_sizes_of(const orc_problem* p, orc_sizes* s)
{
    return guarded([&] {
        Controller ct;
        ct.setup(*p);
        s->X = ct.ps.fullXDim; s->nU = ct.ps.fullUDim; s->nvar = ct.nvar;
        s->meq = 0; s->mineq = 0;
        for (auto* e : ct.eq) s->meq += e->nrConstr;
        for (auto* e : ct.ineq) s->mineq += e->nrConstr;
    });
}

int orc_condense(int nx, int nu, int N, const double* A, const double* B, const double* d, double* Phi,
    double* Psi, double* xi)
{
    return guarded([&] {
        PreviewSystem ps;
        std::vector<double> x0(nx, 0.0);
        ps.system(nx, nu, N, A, B, d, x0.data());
        ps.updateSystem();
        copyOut(Phi, ps.Phi.a); copyOut(Psi, ps.Psi.a); copyOut(xi, ps.xi);
    });
}

int orc_lmpc(const orc_problem* p, orc_outputs* o)
{
    return guarded([&] {
        using clk = std::chrono::steady_clock;
        Controller ct;
        ct.setup(*p);
        auto t0 = clk::now();
        ct.build();
        auto t1 = clk::now();
        bool wantSolve = o->x || o->control || o->trajectory || o->iact || o->fail || o->iter;
        double ts = 0;
        if (wantSolve) {
            auto s0 = clk::now();
            ct.solve();
            ts = std::chrono::duration<double>(clk::now() - s0).count();
        }
        copyOut(o->Phi, ct.ps.Phi.a); copyOut(o->Psi, ct.ps.Psi.a); copyOut(o->xi, ct.ps.xi);
        copyOut(o->Q, ct.Q.a); copyOut(o->c, ct.c);
        copyOut(o->Aeq, ct.Aeq.a); copyOut(o->beq, ct.beq); copyOut(o->Aineq, ct.Aineq.a); copyOut(o->bineq, ct.bineq);
        copyOut(o->lb, ct.lb); copyOut(o->ub, ct.ub);
        if (wantSolve) {
            copyOut(o->x, ct.result); copyOut(o->control, ct.control); copyOut(o->trajectory, ct.trajectory);
            copyOut(o->iact, ct.iact); copyOut(o->lagr, ct.lagr);
            if (o->nact) *o->nact = ct.nact;
            if (o->iter) { o->iter[0] = ct.iter[0]; o->iter[1] = ct.iter[1]; }
            if (o->fail) *o->fail = ct.fail;
            if (o->crval) *o->crval = ct.crval;
        }
        if (o->t_build) *o->t_build = std::chrono::duration<double>(t1 - t0).count();
        if (o->t_solve) *o->t_solve = ts;
    });
}

int orc_quadprog(int n, int meq, int m, const double* Q, const double* c, const double* Aeq, const double* beq,
    const double* Aineq, const double* bineq, const double* lb, const double* ub, double* x, int* iact,
    int* nact, int* iter, double* lagr, double* crval)
{
    int rc = -3;
    int g = guarded([&] { rc = quadprogSolve(n, meq, m, Q, c, Aeq, beq, Aineq, bineq, lb, ub, x, iact, nact, iter, lagr, crval); });
    return g == 0 ? rc : g;
}

double orc_lmpc_batch(const orc_problem* probs, int batch, int threads, double* control, double* trajectory,
    int* fail, int* iter, int* nact, int* iact, double* t_inst)
{
    using clk = std::chrono::steady_clock;
    if (threads < 1) threads = 1;
    std::atomic<int> next(0), bad(0);
    auto t0 = clk::now();
    auto worker = [&] {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= batch) break;
            try {
                auto s0 = clk::now();
                Controller ct;
                ct.setup(probs[i]);
                ct.build();
                ct.solve();
                const int nU = ct.ps.fullUDim, X = ct.ps.fullXDim, q = ct.nrEq + ct.nrIneq + 2 * ct.nvar;
                if (control && !ct.control.empty()) std::memcpy(control + size_t(i) * nU, ct.control.data(), sizeof(double) * nU);
                if (trajectory && !ct.trajectory.empty()) std::memcpy(trajectory + size_t(i) * X, ct.trajectory.data(), sizeof(double) * X);
                if (fail) fail[i] = ct.fail;
                if (iter) { iter[2 * i] = ct.iter[0]; iter[2 * i + 1] = ct.iter[1]; }
                if (nact) nact[i] = ct.nact;
                if (iact) std::memcpy(iact + size_t(i) * q, ct.iact.data(), sizeof(int) * q);
                if (t_inst) t_inst[i] = std::chrono::duration<double>(clk::now() - s0).count();
            } catch (...) { bad.fetch_add(1); }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    double wall = std::chrono::duration<double>(clk::now() - t0).count();
    return bad.load() ? -1.0 : wall;
}

} // extern "C"
