"""TEST INFRASTRUCTURE -- a SECOND, independent restatement of the QP arithmetic behind copra's QuadProgDenseSolver.

oracle/copra_oracle.cpp restates quadprog's Fortran `qpgen2` (+ LINPACK dpofa / dposl / dpori) in C++ scalar loops.  This file
restates the same published algorithm (Goldfarb & Idnani 1983 as coded by B. Turlach in quadprog's solve.QP.f; the
transcription the survey validated) a second time, in numpy, sharing no code with the C++ oracle:
  * numpy's Cholesky + a triangular solve instead of the dpofa / dpori loops,
  * a dense upper-triangular R instead of the packed work array,
  * whole-column Givens rotations instead of scalar loops.
tests/test_oracle.py cross-checks the two on hundreds of random QPs (equalities, mixed finite / inf / DBL_MAX bounds,
hundreds of constraint drops): same x, same active set IN THE SAME ORDER, same iteration and drop counts, same fail
codes.  Parity is still "unpinned" against the compiled eigen-quadprog (the dependency is absent from the reference tree:
`eigen-quadprog`, unpinned, CMakeLists.txt:64); two independent restatements that agree step for step pin the ALGORITHM.

Reference call sites: src/QuadProgSolver.cpp:45-72 (bounds -> 2n dense rows, problem(n, neq, nineq + 2n), solve),
include/QuadProgSolver.h:21-27 (fail codes 0 / 1 / 2).  Only tests/ may import this module.
"""
import numpy as np


def _vsmall():
    v = 1.0e-60
    while True:
        v += v
        if 1.0 + 0.1 * v > 1.0 and 1.0 + 0.2 * v > 1.0:
            return v


def _givens(a, b):
    """quadprog's rotation of the pair (a, b) -> (rho, 0): returns (gc, gs, rho) with rho carrying the sign of a"""
    big, small = max(abs(a), abs(b)), min(abs(a), abs(b))
    rho = np.copysign(big * np.sqrt(1.0 + (small / big) ** 2), a)
    return a / rho, b / rho, rho


def _rotate_cols(M, k, gc, gs, rows=slice(None)):
    """apply the (gc, gs) rotation to columns k, k+1 of M exactly as qpgen2 does (nu form)"""
    nu = gs / (1.0 + gc)
    t = gc * M[rows, k] + gs * M[rows, k + 1]
    M[rows, k + 1] = nu * (M[rows, k] + t) - M[rows, k + 1]
    M[rows, k] = t


def qpgen2(D, dvec, A, b0, meq, max_iter=None):
    """min 1/2 x'Dx - d'x  s.t.  A[:, i]'x = b0[i] (i < meq),  A[:, i]'x >= b0[i] (i >= meq).
    Returns dict(x, ierr, iter=(outer, drops), iact (1-based, add order), nact, u (multipliers in iact order), crval)."""
    D = np.array(D, dtype=np.float64)
    A = np.array(A, dtype=np.float64)      # n x q, modified in place by the equality sign flips
    b0 = np.array(b0, dtype=np.float64)
    dvec = np.asarray(dvec, dtype=np.float64)
    n, q = D.shape[0], A.shape[1]
    vsmall = _vsmall()
    out = dict(x=np.zeros(n), ierr=0, iter=(0, 0), iact=[], nact=0, u=np.zeros(0), crval=0.0)
    try:
        L = np.linalg.cholesky(D)
    except np.linalg.LinAlgError:
        out["ierr"] = 2
        return out
    R0 = L.T                                   # D = R0' R0, upper
    J = np.linalg.solve(R0, np.eye(n))         # R0^-1, upper triangular
    J = np.triu(J)
    x = J @ (J.T @ dvec)                       # D^-1 d
    crval = -0.5 * float(dvec @ x)
    nbv = np.sqrt((A * A).sum(axis=0))         # column norms
    Rm = np.zeros((n, n))                      # factor of the active normals (upper triangular, nact x nact)
    u = np.zeros(n + 2)
    iact = []
    nact, it_outer, it_drop = 0, 0, 0
    cap = max_iter or (50 * q + 100)

    def slack(i):
        s = float(A[:, i] @ x) - b0[i]
        return s

    def drop(it1):
        nonlocal nact, it_drop
        k = it1
        while k < nact - 1:
            if Rm[k + 1, k + 1] != 0.0:
                gc, gs, rho = _givens(Rm[k, k + 1], Rm[k + 1, k + 1])
                if gc != 1.0:
                    cols = slice(k + 1, nact)
                    if gc == 0.0:
                        Rm[[k, k + 1], cols] = Rm[[k + 1, k], cols]
                        J[:, [k, k + 1]] = J[:, [k + 1, k]]
                    else:
                        nu = gs / (1.0 + gc)
                        t = gc * Rm[k, cols] + gs * Rm[k + 1, cols]
                        Rm[k + 1, cols] = nu * (Rm[k, cols] + t) - Rm[k + 1, cols]
                        Rm[k, cols] = t
                        _rotate_cols(J, k, gc, gs)
            Rm[:k + 1, k] = Rm[:k + 1, k + 1]  # column k+1 moves into column k
            u[k] = u[k + 1]
            iact[k] = iact[k + 1]
            k += 1
        u[nact - 1] = u[nact]
        u[nact] = 0.0
        iact.pop()
        Rm[:, nact - 1] = 0.0
        nact -= 1
        it_drop += 1

    while True:
        it_outer += 1
        if it_outer > cap:
            out["ierr"] = 3
            break
        sv = np.empty(q)
        for i in range(q):
            s = slack(i)
            if abs(s) < vsmall:
                s = 0.0
            if i >= meq:
                sv[i] = s
            else:
                sv[i] = -abs(s)
                if s > 0.0:
                    A[:, i] = -A[:, i]
                    b0[i] = -b0[i]
        for i in iact:
            sv[i] = 0.0
        nvl, temp = -1, 0.0
        for i in range(q):
            if sv[i] < temp * nbv[i]:
                nvl, temp = i, sv[i] / nbv[i]
        if nvl < 0:
            break
        s_nvl = sv[nvl]
        failed = False
        while True:  # label 55
            a = A[:, nvl]
            d = J.T @ a
            z = J[:, nact:] @ d[nact:]
            r = np.zeros(nact)
            for i in range(nact - 1, -1, -1):  # back substitution on the upper-triangular R
                r[i] = (d[i] - float(Rm[i, i + 1:nact] @ r[i + 1:])) / Rm[i, i]
            t1inf, t1, it1 = True, 0.0, -1
            for i in range(nact):
                if iact[i] >= meq and r[i] > 0.0:
                    tt_ = u[i] / r[i]
                    if t1inf or tt_ < t1:
                        t1inf, t1, it1 = False, tt_, i
            if abs(float(z @ z)) <= vsmall:
                if t1inf:
                    out["ierr"] = 1
                    failed = True
                    break
                u[:nact] -= t1 * r
                u[nact] += t1
                drop(it1)
                continue
            sm = float(z @ a)
            tt = -s_nvl / sm
            t2min = True
            if not t1inf and t1 < tt:
                tt, t2min = t1, False
            x = x + tt * z
            crval += tt * sm * (tt / 2.0 + u[nact])
            u[:nact] -= tt * r
            u[nact] += tt
            if t2min:
                nact += 1
                iact.append(nvl)
                Rm[:nact - 1, nact - 1] = d[:nact - 1]
                if nact < n:
                    for i in range(n - 1, nact - 1, -1):  # fold d[nact-1:] into its first entry
                        if d[i] == 0.0:
                            continue
                        gc, gs, rho = _givens(d[i - 1], d[i])
                        if gc == 1.0:
                            continue
                        if gc == 0.0:
                            d[i - 1] = gs * rho
                            J[:, [i - 1, i]] = J[:, [i, i - 1]]
                        else:
                            d[i - 1] = rho
                            _rotate_cols(J, i - 1, gc, gs)
                Rm[nact - 1, nact - 1] = d[nact - 1]
                break
            # partial step
            s = slack(nvl)
            if nvl >= meq:
                s_nvl = s
            else:
                s_nvl = -abs(s)
                if s > 0.0:
                    A[:, nvl] = -A[:, nvl]
                    b0[nvl] = -b0[nvl]
            drop(it1)
        if failed:
            break
    out.update(x=x, iter=(it_outer, it_drop), iact=[i + 1 for i in iact], nact=nact, u=u[:nact].copy(), crval=crval)
    return out


def solve_copra_qp(Q, c, Aeq, beq, Aineq, bineq, lb, ub):
    """QuadProgDenseSolver::SI_solve (src/QuadProgSolver.cpp:54-72): the bounds become 2n dense rows [I; -I] appended to the
    inequalities, then eigen-quadprog's mapping onto qpgen2: d = -c, A = [Aeq', -ineqMat'], b0 = [beq; -ineqVec]."""
    Q = np.asarray(Q, dtype=np.float64)
    n = Q.shape[0]
    Aeq = np.zeros((0, n)) if Aeq is None else np.asarray(Aeq, dtype=np.float64).reshape(-1, n)
    Aineq = np.zeros((0, n)) if Aineq is None else np.asarray(Aineq, dtype=np.float64).reshape(-1, n)
    beq = np.zeros(0) if beq is None else np.asarray(beq, dtype=np.float64).reshape(-1)
    bineq = np.zeros(0) if bineq is None else np.asarray(bineq, dtype=np.float64).reshape(-1)
    ineq_mat = np.vstack([Aineq, np.eye(n), -np.eye(n)])
    ineq_vec = np.concatenate([bineq, np.asarray(ub, dtype=np.float64), -np.asarray(lb, dtype=np.float64)])
    A = np.hstack([Aeq.T, -ineq_mat.T])
    b0 = np.concatenate([beq, -ineq_vec])
    r = qpgen2(Q, -np.asarray(c, dtype=np.float64), A, b0, Aeq.shape[0])
    r["fail"] = r["ierr"]
    return r
