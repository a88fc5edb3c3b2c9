"""Parity metrics shared by the CPU and GPU tests."""
import numpy as np


def rel_err(got, ref):
    """norm-wise relative error over finite entries; non-finite / DBL_MAX entries must match exactly."""
    got, ref = np.asarray(got, float), np.asarray(ref, float)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    fin = np.isfinite(ref) & (np.abs(ref) < 1e300)
    assert np.array_equal(got[~fin], ref[~fin]), "infinite / DBL_MAX entries differ"
    if not fin.any():
        return 0.0
    return float(np.abs(got[fin] - ref[fin]).max() / max(np.abs(ref[fin]).max(), 1e-300))


def zeros_preserved(got, ref):
    """structural zeros of the oracle must be exact zeros"""
    got, ref = np.asarray(got), np.asarray(ref)
    return bool(np.all(got[ref == 0.0] == 0.0))


def x_err(got, ref):
    got, ref = np.asarray(got, float), np.asarray(ref, float)
    return float(np.abs(got - ref).max() / max(1.0, np.abs(ref).max()))


def active_set(iact, nact=None):
    iact = np.asarray(iact)
    if nact is not None:
        iact = iact[:nact]
    return set(int(i) for i in iact if i > 0)


def active_set_excused(got, o):
    """active sets are compared as SETS of 1-based indices in the [eq | ineq | upper | lower] space.  A row
    that is in exactly one of the two sets is excused only when it is WEAKLY active at the oracle's solution -- both its
    slack and its multiplier vanish (|s_i| < 1e-9 and u_i < 1e-9, scaled) -- because then the two sets describe the same
    optimum and membership is decided by last-bit rounding.  Returns the number of excused rows; raises otherwise."""
    ref = active_set(o["iact"])
    diff = set(got) ^ ref
    if not diff:
        return 0
    x = np.asarray(o["x"], float)
    n = x.shape[0]
    Aeq, Aineq = np.asarray(o["Aeq"], float).reshape(-1, n), np.asarray(o["Aineq"], float).reshape(-1, n)
    G = np.vstack([Aeq, Aineq, np.eye(n), -np.eye(n)])
    with np.errstate(invalid="ignore"):
        h = np.concatenate([np.asarray(o["beq"], float), np.asarray(o["bineq"], float), np.asarray(o["ub"], float),
                            -np.asarray(o["lb"], float)])
    grad = np.asarray(o["Q"], float) @ x + np.asarray(o["c"], float)
    s_scale, u_scale = max(1.0, np.abs(x).max()), max(1.0, np.abs(grad).max())
    lag = np.asarray(o["lagr"], float)
    for i in sorted(diff):
        k = i - 1
        slack = h[k] - G[k] @ x
        assert abs(slack) < 1e-9 * s_scale and abs(lag[k]) < 1e-9 * u_scale, \
            "active sets differ at row %d which is not weakly active (slack %.3e, multiplier %.3e)" % (i, slack, lag[k])
    return len(diff)


def exact_active_set_solution(o):
    """x of the equality-constrained QP on the oracle's FINAL active set, solved through the KKT system with iterative
    refinement in extended precision.  Used where the oracle itself has drifted: qpgen2 updates its factors in place, and after
    thousands of adds / drops (C5's heaviest instance: 2519 + 1982) its own x is a few 1e-6 away from the exact solution of
    the active set it reports.  There parity means: same active set, same counts, and x within 1e-8 of THIS solution."""
    Q, c = np.asarray(o["Q"], float), np.asarray(o["c"], float)
    n = Q.shape[0]
    Aeq, Aineq = np.asarray(o["Aeq"], float).reshape(-1, n), np.asarray(o["Aineq"], float).reshape(-1, n)
    G = np.vstack([Aeq, Aineq, np.eye(n), -np.eye(n)])
    h = np.concatenate([np.asarray(o["beq"], float), np.asarray(o["bineq"], float), np.asarray(o["ub"], float),
                        -np.asarray(o["lb"], float)])
    act = np.array(sorted(active_set(o["iact"])), dtype=int) - 1
    Na, ba = G[act], h[act]
    k = len(act)
    K = np.block([[Q, Na.T], [Na, np.zeros((k, k))]])
    rhs = np.concatenate([-c, ba])
    Kl, rl = K.astype(np.longdouble), rhs.astype(np.longdouble)
    sol = np.linalg.solve(K, rhs)
    for _ in range(3):
        res = (rl - Kl @ sol.astype(np.longdouble)).astype(float)
        sol = sol + np.linalg.solve(K, res)
    return sol[:n]
