"""Independent numpy restatement of K1-K5 (condense + QP assembly) and a solver-independent KKT
checker.  TEST INFRASTRUCTURE ONLY (cross-checks oracle/copra_oracle.cpp; never used by the product).

It is deliberately written differently from the C++ oracle (closed forms with matrix powers and
Kronecker products instead of the reference's recursions) so that the two restatements do not share
bugs.  Reference lines are cited per function (paths relative to /root/reference).
"""
import numpy as np

DBL_MAX = np.finfo(np.float64).max


def condense(A, B, d, N):
    """src/PreviewSystem.cpp:51-74 in closed form: Phi_i = A^i, Psi_ij = A^(i-1-j) B, xi_i = sum_k<i A^k d."""
    A = np.asarray(A, float)
    B = np.atleast_2d(np.asarray(B, float))
    nx, nu = B.shape
    pw = [np.eye(nx)]
    for _ in range(N):
        pw.append(A @ pw[-1])
    Phi = np.vstack(pw)
    Psi = np.zeros((nx * (N + 1), nu * N))
    for i in range(1, N + 1):
        for j in range(i):
            Psi[i * nx:(i + 1) * nx, j * nu:(j + 1) * nu] = pw[i - 1 - j] @ B
    xi = np.zeros(nx * (N + 1))
    acc = np.zeros(nx)
    for i in range(1, N + 1):
        acc = acc + pw[i - 1] @ d
        xi[i * nx:(i + 1) * nx] = acc
    return Phi, Psi, xi


def _weights(w, rows):
    w = np.ones(rows) if w is None else np.atleast_1d(np.asarray(w, float))
    if w.shape[0] == rows:
        return w
    assert rows % w.shape[0] == 0
    return np.tile(w, rows // w.shape[0])


def cost_terms(cost, nx, nu, N, Phi, Psi, xi):
    """Q, E, f of one step-size or full-size cost (src/costFunctions.cpp:63-215) as ONE stacked
    weighted least-squares term  (T U + Tphi x0 + res)' W (T U + Tphi x0 + res)."""
    X, n = nx * (N + 1), nu * N
    kind = cost["kind"]
    p = np.asarray(cost["p"], float)
    M = None if cost.get("M") is None else np.atleast_2d(np.asarray(cost["M"], float))
    Nm = None if cost.get("N") is None else np.atleast_2d(np.asarray(cost["N"], float))
    if kind == "trajectory":
        if M.shape[1] == nx:
            Mf, pf, wf = np.kron(np.eye(N + 1), M), np.tile(p, N + 1), np.tile(_weights(cost.get("w"), len(p)), N + 1)
        else:
            Mf, pf, wf = M, p, _weights(cost.get("w"), len(p))
        T, Tphi, res = Mf @ Psi, Mf @ Phi, Mf @ xi - pf
    elif kind == "target":
        Mf = np.hstack([np.zeros((M.shape[0], X - nx)), M])
        wf = _weights(cost.get("w"), len(p))
        T, Tphi, res = Mf @ Psi, Mf @ Phi, Mf @ xi - p
    elif kind == "control":
        if Nm.shape[1] == nu:
            Nf, pf, wf = np.kron(np.eye(N), Nm), np.tile(p, N), np.tile(_weights(cost.get("w"), len(p)), N)
        else:
            Nf, pf, wf = Nm, p, _weights(cost.get("w"), len(p))
        T, Tphi, res = Nf, np.zeros((Nf.shape[0], nx)), -pf
    elif kind == "mixed":
        if M.shape[1] == nx:
            Mf = np.hstack([np.kron(np.eye(N), M), np.zeros((M.shape[0] * N, nx))])  # x_N not penalised
            Nf, pf, wf = np.kron(np.eye(N), Nm), np.tile(p, N), np.tile(_weights(cost.get("w"), len(p)), N)
        else:
            Mf, Nf, pf, wf = M, Nm, p, _weights(cost.get("w"), len(p))
        T, Tphi, res = Mf @ Psi + Nf, Mf @ Phi, Mf @ xi - pf
    else:
        raise ValueError(kind)
    W = np.diag(wf)
    return T.T @ W @ T, Tphi.T @ W @ T, res @ W @ T


def constraint_rows(cstr, nx, nu, N, Phi, Psi, xi):
    """A, Y, z of one constraint (src/constraints.cpp:66-315):  Y x0 + A U (<=|==) z."""
    kind = cstr["kind"]
    if kind == "trajectory_bound":
        lo, up = np.asarray(cstr["lower"], float), np.asarray(cstr["upper"], float)
        rows, rhs = [], []
        for bound, lines in ((lo, np.nonzero(lo != -np.inf)[0]), (up, np.nonzero(up != np.inf)[0])):
            steps = range(N + 1) if lo.shape[0] == nx else range(1)
            for s in steps:
                for ln in lines:
                    rows.append(ln + nx * s)
                    rhs.append(bound[ln])
        rows = np.asarray(rows, int)
        # NB (quirk Q1): lower rows are NOT negated by the reference
        return Psi[rows], Phi[rows], np.asarray(rhs) - xi[rows]
    f = np.asarray(cstr["f"], float)
    E = None if cstr.get("E") is None else np.atleast_2d(np.asarray(cstr["E"], float))
    G = None if cstr.get("G") is None else np.atleast_2d(np.asarray(cstr["G"], float))
    if kind == "trajectory":
        Ef, ff = (np.kron(np.eye(N + 1), E), np.tile(f, N + 1)) if E.shape[1] == nx else (E, f)
        return Ef @ Psi, Ef @ Phi, ff - Ef @ xi
    if kind == "control":
        Gf, ff = (np.kron(np.eye(N), G), np.tile(f, N)) if G.shape[1] == nu else (G, f)
        return Gf, np.zeros((Gf.shape[0], nx)), ff
    if kind == "mixed":
        if E.shape[1] == nx:
            Ef = np.hstack([np.kron(np.eye(N), E), np.zeros((E.shape[0] * N, nx))])
            Gf, ff = np.kron(np.eye(N), G), np.tile(f, N)
        else:
            Ef, Gf, ff = E, G, f
        return Ef @ Psi + Gf, Ef @ Phi, ff - Ef @ xi
    raise ValueError(kind)


def build_qp(prob):
    """LMPC::updateSystem + makeQPForm (src/LMPC.cpp:225-280) or the InitialStateLMPC variant
    (src/InitialStateLMPC.cpp:77-122).  `prob` is a single-instance dict (workloads.instance)."""
    nx, nu, N = prob["nx"], prob["nu"], prob["N"]
    x0 = np.asarray(prob["x0"], float)
    Phi, Psi, xi = condense(prob["A"], prob["B"], prob["d"], N)
    n = nu * N
    ist = bool(prob.get("initial_state"))
    Qs, Es, fs = np.zeros((n, n)), np.zeros((nx, n)), np.zeros(n)
    for c in prob["costs"]:
        q_, e_, f_ = cost_terms(c, nx, nu, N, Phi, Psi, xi)
        Qs, Es, fs = Qs + q_, Es + e_, fs + f_
    Qs = Qs + 1e-6 * np.eye(n)
    eq, ineq = [], []
    lb, ub = np.full(n, -DBL_MAX), np.full(n, DBL_MAX)
    off = 0
    for c in prob["constraints"]:
        if c["kind"] == "control_bound":
            lo, up = np.asarray(c["lower"], float), np.asarray(c["upper"], float)
            reps = 1 if lo.shape[0] == n else N
            lb[off:off + reps * lo.shape[0]] = np.tile(lo, reps)
            ub[off:off + reps * lo.shape[0]] = np.tile(up, reps)
            off += reps * lo.shape[0]
            continue
        A_, Y_, z_ = constraint_rows(c, nx, nu, N, Phi, Psi, xi)
        (ineq if c.get("is_ineq", True) or c["kind"] == "trajectory_bound" else eq).append((A_, Y_, z_))

    def stack(lst):
        if not lst:
            return np.zeros((0, n)), np.zeros((0, nx)), np.zeros(0)
        return np.vstack([t[0] for t in lst]), np.vstack([t[1] for t in lst]), np.concatenate([t[2] for t in lst])

    Ae, Ye, ze = stack(eq)
    Ai, Yi, zi = stack(ineq)
    out = dict(Phi=Phi, Psi=Psi, xi=xi)
    if not ist:
        out.update(Q=Qs, c=Es.T @ x0 + fs, Aeq=Ae, beq=ze - Ye @ x0, Aineq=Ai, bineq=zi - Yi @ x0, lb=lb, ub=ub)
    else:
        R = np.zeros((nx, nx)) if prob.get("R") is None else np.asarray(prob["R"], float)
        r = np.zeros(nx) if prob.get("r") is None else np.asarray(prob["r"], float)
        x0lb = x0 if prob.get("x0lb") is None else np.asarray(prob["x0lb"], float)
        x0ub = x0 if prob.get("x0ub") is None else np.asarray(prob["x0ub"], float)
        H = np.block([[R + Es @ np.linalg.inv(Qs) @ Es.T, Es], [Es.T, Qs]])
        out.update(Q=H, c=np.concatenate([r, fs]), Aeq=np.hstack([Ye, Ae]), beq=ze, Aineq=np.hstack([Yi, Ai]),
                   bineq=zi, lb=np.concatenate([x0lb, lb]), ub=np.concatenate([x0ub, ub]))
    return out


def kkt_residuals(Q, c, Aeq, beq, Aineq, bineq, lb, ub, x, iact):
    """Solver-independent acceptance: given x and the active set (1-based indices in
    [eq | ineq | upper | lower] space), recover multipliers by least squares and report
    stationarity / primal / dual / complementarity residuals (all scaled)."""
    Q, c, x = np.asarray(Q, float), np.asarray(c, float), np.asarray(x, float)
    n = x.shape[0]
    Aeq = np.zeros((0, n)) if Aeq is None else np.asarray(Aeq, float).reshape(-1, n)
    Aineq = np.zeros((0, n)) if Aineq is None else np.asarray(Aineq, float).reshape(-1, n)
    beq = np.zeros(0) if beq is None else np.asarray(beq, float)
    bineq = np.zeros(0) if bineq is None else np.asarray(bineq, float)
    meq, m = Aeq.shape[0], Aineq.shape[0]
    # all rows in "G x <= h" form, equalities first
    G = np.vstack([Aeq, Aineq, np.eye(n), -np.eye(n)])
    with np.errstate(invalid="ignore"):
        h = np.concatenate([beq, bineq, ub, -np.asarray(lb, float)])
    act = np.asarray(sorted(set(int(i) - 1 for i in iact)), int)
    act = np.union1d(act, np.arange(meq)).astype(int)
    grad = Q @ x + c
    scale = max(1.0, np.abs(grad).max(), np.abs(c).max())
    if act.size:
        Ga = G[act]
        lam, *_ = np.linalg.lstsq(Ga.T, -grad, rcond=None)
        stat = np.abs(grad + Ga.T @ lam).max()
    else:
        lam, stat = np.zeros(0), np.abs(grad).max()
    fin = np.isfinite(h) & (np.abs(h) < 1e300)
    slack = np.where(fin, h - G @ x, np.inf)
    xs = max(1.0, np.abs(x).max())
    prim_ineq = max(0.0, -(slack[meq:].min())) if slack[meq:].size else 0.0
    prim_eq = np.abs(slack[:meq]).max() if meq else 0.0
    is_ineq_act = act >= meq
    dual = max(0.0, -(lam[is_ineq_act].min())) if is_ineq_act.any() else 0.0
    comp = np.abs(lam[is_ineq_act] * slack[act[is_ineq_act]]).max() if is_ineq_act.any() else 0.0
    return dict(stationarity=stat / scale, primal=max(prim_ineq, prim_eq) / xs, dual=dual / scale,
                complementarity=comp / scale, multipliers=lam, active=act + 1)
