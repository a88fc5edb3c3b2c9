"""ctypes front-end of the CPU ORACLE (oracle/copra_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (copra_b200) never imports this module.

Matrices cross this boundary as numpy arrays in *logical* (rows, cols) shape; they are converted
to column-major (Eigen layout, ld == rows) on the way in and back on the way out.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libcopra_oracle.so")

COST_KINDS = {"trajectory": 0, "target": 1, "control": 2, "mixed": 3}
CSTR_KINDS = {"trajectory": 0, "control": 1, "mixed": 2, "trajectory_bound": 3, "control_bound": 4}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class OrcCost(C.Structure):
    _fields_ = [("kind", C.c_int), ("rows", C.c_int), ("colsM", C.c_int), ("colsN", C.c_int),
                ("wrows", C.c_int), ("autospan", C.c_int),
                ("M", _dp), ("N", _dp), ("p", _dp), ("w", _dp)]


class OrcConstraint(C.Structure):
    _fields_ = [("kind", C.c_int), ("rows", C.c_int), ("colsE", C.c_int), ("colsG", C.c_int),
                ("is_ineq", C.c_int), ("autospan", C.c_int),
                ("E", _dp), ("G", _dp), ("f", _dp), ("lower", _dp), ("upper", _dp)]


class OrcProblem(C.Structure):
    _fields_ = [("nx", C.c_int), ("nu", C.c_int), ("N", C.c_int),
                ("A", _dp), ("B", _dp), ("d", _dp), ("x0", _dp),
                ("ncost", C.c_int), ("costs", C.POINTER(OrcCost)),
                ("ncstr", C.c_int), ("cstrs", C.POINTER(OrcConstraint)),
                ("initial_state", C.c_int),
                ("R", _dp), ("r", _dp), ("x0lb", _dp), ("x0ub", _dp)]


class OrcSizes(C.Structure):
    _fields_ = [("X", C.c_int), ("nU", C.c_int), ("nvar", C.c_int), ("meq", C.c_int), ("mineq", C.c_int)]


class OrcOutputs(C.Structure):
    _fields_ = [(k, _dp) for k in ("Phi", "Psi", "xi", "Q", "c", "Aeq", "beq", "Aineq", "bineq",
                                   "lb", "ub", "x", "control", "trajectory")] + \
               [("iact", _ip), ("nact", _ip), ("iter", _ip), ("fail", _ip),
                ("lagr", _dp), ("crval", _dp), ("t_build", _dp), ("t_solve", _dp)]


def build(force=False):
    """Compile libcopra_oracle.so with the committed Makefile (g++ only)."""
    src = os.path.join(_HERE, "copra_oracle.cpp")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libcopra_oracle.so"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        try:
            _lib = C.CDLL(_LIB)
        except OSError:
            build(force=True)
            _lib = C.CDLL(_LIB)
        _lib.orc_last_error.restype = C.c_char_p
        _lib.orc_lmpc_batch.restype = C.c_double
        _lib.orc_lmpc_batch.argtypes = [C.POINTER(OrcProblem), C.c_int, C.c_int, _dp, _dp, _ip, _ip, _ip, _ip, _dp]
        _lib.orc_condense.argtypes = [C.c_int] * 3 + [_dp] * 6
        _lib.orc_quadprog.argtypes = [C.c_int] * 3 + [_dp] * 8 + [_dp, _ip, _ip, _ip, _dp, _dp]
    return _lib


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__("oracle error %d: %s" % (code, msg))
        self.code = code  # -1 std::domain_error, -2 std::runtime_error


def _check(rc):
    if rc != 0:
        raise OracleError(rc, lib().orc_last_error().decode())


def _cm(a):
    """logical (r, c) array -> flat column-major float64 buffer."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        return np.ascontiguousarray(a)
    return np.ascontiguousarray(a.T).reshape(-1)


def _ptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _from_cm(buf, r, c):
    return np.asarray(buf).reshape(c, r).T.copy()


class _Keep(list):
    """keeps numpy buffers alive while a ctypes struct points at them"""

    def cm(self, a):
        if a is None:
            return None
        b = _cm(a)
        self.append(b)
        return _ptr(b)


def make_problem(prob, keep):
    """prob: dict describing ONE instance (see copra_b200.workloads.instance()).
    Returns an OrcProblem whose pointers are kept alive by `keep`."""
    p = OrcProblem()
    p.nx, p.nu, p.N = int(prob["nx"]), int(prob["nu"]), int(prob["N"])
    p.A, p.B, p.d, p.x0 = keep.cm(prob["A"]), keep.cm(prob["B"]), keep.cm(prob["d"]), keep.cm(prob["x0"])
    costs = (OrcCost * max(1, len(prob["costs"])))()
    for i, c in enumerate(prob["costs"]):
        oc = costs[i]
        oc.kind = COST_KINDS[c["kind"]]
        pvec = np.asarray(c["p"], dtype=np.float64)
        oc.rows = int(pvec.shape[0])
        if c.get("M") is not None:
            M = np.atleast_2d(np.asarray(c["M"], dtype=np.float64))
            oc.colsM = M.shape[1]
            oc.rows = M.shape[0]
            oc.M = keep.cm(M)
        if c.get("N") is not None:
            Nm = np.atleast_2d(np.asarray(c["N"], dtype=np.float64))
            oc.colsN = Nm.shape[1]
            oc.rows = Nm.shape[0]
            oc.N = keep.cm(Nm)
        oc.p = keep.cm(pvec)
        w = c.get("w")
        if w is not None:
            w = np.atleast_1d(np.asarray(w, dtype=np.float64))
            oc.w = keep.cm(w)
            oc.wrows = w.shape[0]
        oc.autospan = int(bool(c.get("autospan", False)))
    keep.append(costs)
    p.ncost, p.costs = len(prob["costs"]), costs
    cstrs = (OrcConstraint * max(1, len(prob["constraints"])))()
    for i, c in enumerate(prob["constraints"]):
        oc = cstrs[i]
        oc.kind = CSTR_KINDS[c["kind"]]
        oc.is_ineq = int(bool(c.get("is_ineq", True)))
        oc.autospan = int(bool(c.get("autospan", False)))
        if c["kind"] in ("trajectory_bound", "control_bound"):
            lo = np.asarray(c["lower"], dtype=np.float64)
            oc.rows = lo.shape[0]
            oc.lower, oc.upper = keep.cm(lo), keep.cm(c["upper"])
            if np.asarray(c["upper"]).shape[0] != oc.rows:
                raise OracleError(-1, "lower/upper rows")
        else:
            f = np.asarray(c["f"], dtype=np.float64)
            oc.rows = f.shape[0]
            oc.f = keep.cm(f)
            if c.get("E") is not None:
                E = np.atleast_2d(np.asarray(c["E"], dtype=np.float64))
                oc.colsE, oc.E = E.shape[1], keep.cm(E)
                if E.shape[0] != oc.rows:
                    raise OracleError(-1, "E/f rows")
            if c.get("G") is not None:
                G = np.atleast_2d(np.asarray(c["G"], dtype=np.float64))
                oc.colsG, oc.G = G.shape[1], keep.cm(G)
                if G.shape[0] != oc.rows:
                    raise OracleError(-1, "G/f rows")
    keep.append(cstrs)
    p.ncstr, p.cstrs = len(prob["constraints"]), cstrs
    p.initial_state = int(bool(prob.get("initial_state", False)))
    for k in ("R", "r", "x0lb", "x0ub"):
        if prob.get(k) is not None:
            setattr(p, k, keep.cm(prob[k]))
    return p


def sizes(prob):
    keep = _Keep()
    p = make_problem(prob, keep)
    s = OrcSizes()
    _check(lib().orc_sizes_of(C.byref(p), C.byref(s)))
    return dict(X=s.X, nU=s.nU, nvar=s.nvar, meq=s.meq, mineq=s.mineq)


def condense(A, B, d, N):
    A = np.asarray(A, dtype=np.float64)
    B = np.atleast_2d(np.asarray(B, dtype=np.float64))
    nx, nu = A.shape[0], B.shape[1]
    X, n = nx * (N + 1), nu * N
    Phi, Psi, xi = np.zeros(X * nx), np.zeros(X * n), np.zeros(X)
    a, b, dd = _cm(A), _cm(B), _cm(d)
    _check(lib().orc_condense(nx, nu, N, _ptr(a), _ptr(b), _ptr(dd), _ptr(Phi), _ptr(Psi), _ptr(xi)))
    return _from_cm(Phi, X, nx), _from_cm(Psi, X, n), xi


def lmpc(prob, solve=True):
    """Build (and solve) one LMPC / InitialStateLMPC instance.  Returns a dict of every stage."""
    keep = _Keep()
    p = make_problem(prob, keep)
    s = OrcSizes()
    _check(lib().orc_sizes_of(C.byref(p), C.byref(s)))
    X, nU, nv, meq, m, nx = s.X, s.nU, s.nvar, s.meq, s.mineq, p.nx
    q = meq + m + 2 * nv
    bufs = dict(Phi=np.zeros(X * nx), Psi=np.zeros(X * nU), xi=np.zeros(X), Q=np.zeros(nv * nv), c=np.zeros(nv),
                Aeq=np.zeros(meq * nv), beq=np.zeros(meq), Aineq=np.zeros(m * nv), bineq=np.zeros(m),
                lb=np.zeros(nv), ub=np.zeros(nv))
    if solve:
        bufs.update(x=np.zeros(nv), control=np.zeros(nU), trajectory=np.zeros(X), lagr=np.zeros(q))
    o = OrcOutputs()
    for k, v in bufs.items():
        setattr(o, k, _ptr(v))
    iact = np.zeros(q, dtype=np.int32)
    nact, fail = C.c_int(0), C.c_int(-1)
    it = np.zeros(2, dtype=np.int32)
    crval, tb, ts = C.c_double(0), C.c_double(0), C.c_double(0)
    if solve:
        o.iact, o.nact, o.iter, o.fail = iact.ctypes.data_as(_ip), C.pointer(nact), it.ctypes.data_as(_ip), C.pointer(fail)
        o.crval = C.pointer(crval)
    o.t_build, o.t_solve = C.pointer(tb), C.pointer(ts)
    _check(lib().orc_lmpc(C.byref(p), C.byref(o)))
    out = dict(sizes=dict(X=X, nU=nU, nvar=nv, meq=meq, mineq=m),
               Phi=_from_cm(bufs["Phi"], X, nx), Psi=_from_cm(bufs["Psi"], X, nU), xi=bufs["xi"],
               Q=_from_cm(bufs["Q"], nv, nv), c=bufs["c"], Aeq=_from_cm(bufs["Aeq"], meq, nv), beq=bufs["beq"],
               Aineq=_from_cm(bufs["Aineq"], m, nv), bineq=bufs["bineq"], lb=bufs["lb"], ub=bufs["ub"],
               t_build=tb.value, t_solve=ts.value)
    if solve:
        out.update(x=bufs["x"], control=bufs["control"], trajectory=bufs["trajectory"], lagr=bufs["lagr"],
                   iact=iact[:nact.value].copy(), nact=nact.value, iter=(int(it[0]), int(it[1])), fail=fail.value,
                   crval=crval.value)
    return out


def quadprog(Q, c, Aeq, beq, Aineq, bineq, lb, ub):
    """QuadProgDenseSolver::SI_solve on a raw QP; returns dict(x, iact, nact, iter, fail, lagr, crval)."""
    Q = np.asarray(Q, dtype=np.float64)
    n = Q.shape[0]
    Aeq = np.zeros((0, n)) if Aeq is None else np.asarray(Aeq, dtype=np.float64).reshape(-1, n)
    Aineq = np.zeros((0, n)) if Aineq is None else np.asarray(Aineq, dtype=np.float64).reshape(-1, n)
    meq, m = Aeq.shape[0], Aineq.shape[0]
    beq = np.zeros(0) if beq is None else np.asarray(beq, dtype=np.float64)
    bineq = np.zeros(0) if bineq is None else np.asarray(bineq, dtype=np.float64)
    q = meq + m + 2 * n
    x, lagr = np.zeros(n), np.zeros(q)
    iact, it = np.zeros(q, dtype=np.int32), np.zeros(2, dtype=np.int32)
    nact, crval = C.c_int(0), C.c_double(0)
    bufs = [_cm(v) for v in (Q, c, Aeq, beq, Aineq, bineq, lb, ub)]
    fail = lib().orc_quadprog(n, meq, m, *[_ptr(b) for b in bufs], _ptr(x), iact.ctypes.data_as(_ip),
                              C.byref(nact), it.ctypes.data_as(_ip), _ptr(lagr), C.byref(crval))
    if fail < 0:
        _check(fail)
    return dict(x=x, iact=iact[:nact.value].copy(), nact=nact.value, iter=(int(it[0]), int(it[1])), fail=fail,
                lagr=lagr, crval=crval.value)


def lmpc_batch(probs, threads=None):
    """CPU baseline: one instance per host thread.  probs: list of per-instance dicts (same shape)."""
    keep = _Keep()
    arr = (OrcProblem * len(probs))()
    for i, pr in enumerate(probs):
        arr[i] = make_problem(pr, keep)
    s = OrcSizes()
    _check(lib().orc_sizes_of(C.byref(arr[0]), C.byref(s)))
    B = len(probs)
    q = s.meq + s.mineq + 2 * s.nvar
    control, traj = np.zeros((B, s.nU)), np.zeros((B, s.X))
    fail, it, nact = np.zeros(B, np.int32), np.zeros((B, 2), np.int32), np.zeros(B, np.int32)
    iact, tinst = np.zeros((B, q), np.int32), np.zeros(B)
    threads = threads or hw_threads()
    wall = lib().orc_lmpc_batch(arr, B, threads, _ptr(control), _ptr(traj), fail.ctypes.data_as(_ip),
                                it.ctypes.data_as(_ip), nact.ctypes.data_as(_ip), iact.ctypes.data_as(_ip), _ptr(tinst))
    if wall < 0:
        raise OracleError(-3, "an instance threw inside orc_lmpc_batch")
    return dict(wall=wall, threads=threads, control=control, trajectory=traj, fail=fail, iter=it, nact=nact,
                iact=iact, t_inst=tinst)


def hw_threads():
    """host threads this process may use: the scheduler affinity mask, not the machine's core count"""
    import os
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, lib().orc_hw_threads())
