"""pyCopra-compatible front end over the B200 engine.

The reference ships a Boost.Python module `pyCopra` (binding/python/CopraBindings.cpp:86-298) exposing PreviewSystem,
the four costs, the five constraints, AutoSpan, SolverFlag and LMPC.  This module offers the same names, constructor
signatures, method names (`add_cost`, `add_constraint`, `solve`, `control`, `trajectory`, `solve_time`, ...), error
behaviour (dimension errors surface as RuntimeError, like Boost.Python's translation of std::domain_error) and the
use_count based auto-removal of costs / constraints the caller no longer holds (src/LMPC.cpp:288-307) -- so that
`import copra_b200.pycopra as copra` runs the reference's python test scenarios (binding/python/tests/pyTests.py)
unchanged.  Every solve goes through the C ABI (`copra_b200_lmpc_run`, K1..K7 on the GPU); nothing here computes.
InitialStateLMPC, which the reference binding never exposed (CopraBindings.cpp:286 is stale), is available too.
"""
import enum
import sys
import time

import numpy as np

from . import capi

_engine = None


def _eng():
    """process-wide engine handle (one GPU, like the C++ facade's b200::handle())"""
    global _engine
    if _engine is None:
        _engine = capi.Engine(0)
    return _engine


def _domain_error(msg):
    raise RuntimeError(msg)


class SolverFlag(enum.Enum):  # include/solverUtils.h:34-50; every flag resolves to the GPU Goldfarb-Idnani solver
    DEFAULT = 0
    QuadProgDense = 1
    B200 = 2


class ConstraintFlag(enum.Enum):  # include/constraints.h:21-26
    Constraint = 0
    EqualityConstraint = 1
    InequalityConstraint = 2
    BoundConstraint = 3


class AutoSpan:
    """src/AutoSpan.cpp:10-48: block-diagonal / tiled extension to a new row dimension"""

    @staticmethod
    def span_matrix(mat, new_dim, add_cols=0):
        mat = np.atleast_2d(np.asarray(mat, dtype=np.float64))
        rows, cols = mat.shape
        if new_dim == rows:
            return mat.copy()
        steps = new_dim // rows
        if steps * rows != new_dim:
            _domain_error("new dimension %d is not a multiple of the %d rows of the matrix" % (new_dim, rows))
        out = np.zeros((new_dim, cols * (steps + add_cols)))
        for i in range(steps):
            out[i * rows:(i + 1) * rows, i * cols:(i + 1) * cols] = mat
        return out

    @staticmethod
    def span_vector(vec, new_dim):
        vec = np.asarray(vec, dtype=np.float64).reshape(-1)
        rows = vec.shape[0]
        if new_dim == rows:
            return vec.copy()
        steps = new_dim // rows
        if steps * rows != new_dim:
            _domain_error("new dimension %d is not a multiple of the %d rows of the vector" % (new_dim, rows))
        return np.tile(vec, steps)


class PreviewSystem:
    """include/PreviewSystem.h:24-78 / src/PreviewSystem.cpp:17-74"""

    def __init__(self, *args):
        self.is_updated = False
        self.nr_u_Step = self.nr_x_Step = self.x_dim = self.u_dim = self.full_x_dim = self.full_u_dim = 0
        self.x0 = self.A = self.B = self.d = self.Phi = self.Psi = self.xi = None
        if args:
            self.system(*args)

    def system(self, state, control, bias, x_init, number_of_steps):
        A = np.atleast_2d(np.asarray(state, dtype=np.float64))
        B = np.asarray(control, dtype=np.float64)
        B = B.reshape(-1, 1) if B.ndim == 1 else B
        d = np.asarray(bias, dtype=np.float64).reshape(-1)
        x0 = np.asarray(x_init, dtype=np.float64).reshape(-1)
        if number_of_steps <= 0:
            _domain_error("The number of step sould be a positive number! ")
        if A.shape[0] != A.shape[1]:
            _domain_error("state should be a square matrix")
        if x0.shape[0] != A.shape[0] or B.shape[0] != A.shape[0] or d.shape[0] != A.shape[0]:
            _domain_error("xInit, control and bias should have as many rows as state")
        self.A, self.B, self.d, self.x0 = A.copy(), B.copy(), d.copy(), x0.copy()
        self.x_dim, self.u_dim = A.shape[0], B.shape[1]
        self.nr_u_Step, self.nr_x_Step = int(number_of_steps), int(number_of_steps) + 1
        self.full_x_dim, self.full_u_dim = self.x_dim * self.nr_x_Step, self.u_dim * self.nr_u_Step
        self.Phi = np.zeros((self.full_x_dim, self.x_dim))
        self.Psi = np.zeros((self.full_x_dim, self.full_u_dim))
        self.xi = np.zeros(self.full_x_dim)
        self.is_updated = False

    def x_init(self, x0):  # PreviewSystem::xInit, include/PreviewSystem.h:52-54
        x0 = np.asarray(x0, dtype=np.float64).reshape(-1)
        if x0.shape[0] != self.x_dim:
            _domain_error("xInit has a bad dimension")
        self.x0 = x0.copy()

    def update_system(self):  # K1 on the GPU
        Phi, Psi, xi = _eng().condense(self.A, self.B, self.d, self.nr_u_Step, want_psi=True)
        self.Phi, self.Psi, self.xi = Phi[0].copy(), Psi[0].copy(), xi[0].copy()
        self.is_updated = True

    def _base(self):
        return dict(nx=self.x_dim, nu=self.u_dim, N=self.nr_u_Step, batch=1, A=self.A, B=self.B, d=self.d, x0=self.x0)


# ------------------------------------------------------------------------------------------------ costs
class CostFunction:
    """include/costFunctions.h:30-103"""

    _kind = None

    def __init__(self, name, rows):
        self._name = name
        self._w = np.ones(rows)
        self._M = self._N = None
        self._Q = self._c = None

    def name(self):
        return self._name

    def weights(self, w):
        self._w = np.asarray(w, dtype=np.float64).reshape(-1).copy()

    def weight(self, w):
        self._w = np.full(self._p.shape[0], float(w))

    def auto_span(self):
        dims = [self._p.shape[0], self._w.shape[0]] + [m.shape[0] for m in (self._M, self._N) if m is not None]
        md = max(dims)
        if self._M is not None:
            self._M = AutoSpan.span_matrix(self._M, md, 1 if self._kind == "mixed" else 0)
        if self._N is not None:
            self._N = AutoSpan.span_matrix(self._N, md)
        self._p = AutoSpan.span_vector(self._p, md)
        self._w = AutoSpan.span_vector(self._w, md)

    def _desc(self):
        return dict(kind=self._kind, M=self._M, N=self._N, p=self._p, w=self._w)

    def initialize_cost(self, ps):
        """dimension checks of initializeCost (src/costFunctions.cpp:44-61,88-104,122-137,173-193); the C ABI only sees
        pointers, so the shapes are checked here, then the engine validates the description as a whole"""
        rows = self._p.shape[0]
        for nm, mat, dim, full in (("M", self._M, ps.x_dim, ps.full_x_dim), ("N", self._N, ps.u_dim, ps.full_u_dim)):
            if mat is None:
                continue
            if mat.shape[0] != rows:
                _domain_error("%s and p should have the same number of rows (%d vs %d); try auto_span" % (nm, mat.shape[0], rows))
            if mat.shape[1] not in (dim, full):
                _domain_error("%s should have %d or %d columns, it has %d" % (nm, dim, full, mat.shape[1]))
        if self._M is not None and self._N is not None and (self._M.shape[1] == ps.x_dim) != (self._N.shape[1] == ps.u_dim):
            _domain_error("M and N should both be step-size or both be full-size")
        if self._w.shape[0] != rows and (self._w.shape[0] == 0 or rows % self._w.shape[0] != 0):
            _domain_error("weights badly dimensioned")
        _eng().sizes(capi.HostBatch(dict(ps._base(), costs=[self._desc()], constraints=[])))

    def update(self, ps):
        """this cost alone: Q() and c() (K1 + K2 with one cost family, no 1e-6 regulariser)"""
        hb = capi.HostBatch(dict(ps._base(), costs=[self._desc()], constraints=[]))
        hb.problem.flags = capi.FLAG_NO_REG
        eng = _eng()
        eng.lmpc_build(hb)
        self._Q = eng.download(hb, "Q")[0]
        self._c = eng.download(hb, "c")[0]

    def Q(self):
        return self._Q

    def c(self):
        return self._c


class TrajectoryCost(CostFunction):
    _kind = "trajectory"

    def __init__(self, M, p):
        self._p = np.asarray(p, dtype=np.float64).reshape(-1).copy()
        super().__init__("TrajectoryCost", self._p.shape[0])
        self._M = np.atleast_2d(np.asarray(M, dtype=np.float64)).copy()


class TargetCost(CostFunction):
    _kind = "target"

    def __init__(self, M, p):
        self._p = np.asarray(p, dtype=np.float64).reshape(-1).copy()
        super().__init__("TargetCost", self._p.shape[0])
        self._M = np.atleast_2d(np.asarray(M, dtype=np.float64)).copy()

    def auto_span(self):  # src/costFunctions.cpp:84-86: nothing to span
        pass


class ControlCost(CostFunction):
    _kind = "control"

    def __init__(self, N, p):
        self._p = np.asarray(p, dtype=np.float64).reshape(-1).copy()
        super().__init__("ControlCost", self._p.shape[0])
        self._N = np.atleast_2d(np.asarray(N, dtype=np.float64)).copy()


class MixedCost(CostFunction):
    _kind = "mixed"

    def __init__(self, M, N, p):
        self._p = np.asarray(p, dtype=np.float64).reshape(-1).copy()
        super().__init__("MixedCost", self._p.shape[0])
        self._M = np.atleast_2d(np.asarray(M, dtype=np.float64)).copy()
        self._N = np.atleast_2d(np.asarray(N, dtype=np.float64)).copy()


# ------------------------------------------------------------------------------------------- constraints
class Constraint:
    """include/constraints.h:40-99"""

    _kind = None

    def __init__(self, name):
        self._name = name
        self._nr_constr = 0
        self._E = self._G = self._f = self._lower = self._upper = None
        self._A = self._b = None
        self._ineq = True

    def name(self):
        return self._name

    def nr_constr(self):
        return self._nr_constr

    def _desc(self):
        return dict(kind=self._kind, E=self._E, G=self._G, f=self._f, lower=self._lower, upper=self._upper, is_ineq=self._ineq)

    def initialize_constraint(self, ps):
        """dimension checks of initializeConstraint (src/constraints.cpp:45-64,106-135,171-195,240-268,325-345)"""
        if self._f is not None:
            rows = self._f.shape[0]
            for nm, mat, dim, full in (("E", self._E, ps.x_dim, ps.full_x_dim), ("G", self._G, ps.u_dim, ps.full_u_dim)):
                if mat is None:
                    continue
                if mat.shape[0] != rows:
                    _domain_error("%s and f should have the same number of rows (%d vs %d); try auto_span" % (nm, mat.shape[0], rows))
                if mat.shape[1] not in (dim, full):
                    _domain_error("%s should have %d or %d columns, it has %d" % (nm, dim, full, mat.shape[1]))
            if self._E is not None and self._G is not None and (self._E.shape[1] == ps.x_dim) != (self._G.shape[1] == ps.u_dim):
                _domain_error("E and G should both be step-size or both be full-size")
        else:
            dim, full = (ps.x_dim, ps.full_x_dim) if self._kind == "trajectory_bound" else (ps.u_dim, ps.full_u_dim)
            if self._lower.shape[0] not in (dim, full):
                _domain_error("lower / upper should have %d or %d rows, they have %d" % (dim, full, self._lower.shape[0]))
        sz = _eng().sizes(capi.HostBatch(dict(ps._base(), costs=[], constraints=[self._desc()])))
        self._nr_constr = sz["meq"] + sz["mineq"] if self._kind not in ("control_bound",) else sz["nU"]

    def update(self, ps):
        """this constraint alone: A() and b() of its rows"""
        if self._kind == "control_bound":
            return
        hb = capi.HostBatch(dict(ps._base(), costs=[], constraints=[self._desc()]))
        eng = _eng()
        eng.lmpc_build(hb)
        eq = not self._ineq and self._kind != "trajectory_bound"
        self._A = eng.download(hb, "Aeq" if eq else "Aineq")[0]
        self._b = eng.download(hb, "beq" if eq else "bineq")[0]

    def A(self):
        return self._A

    def b(self):
        return self._b


class _EqIneq(Constraint):
    def __init__(self, qualifier, is_inequality):
        super().__init__(qualifier + (" inequality constraint" if is_inequality else " equality constraint"))
        self._ineq = bool(is_inequality)

    def constraint_type(self):
        return ConstraintFlag.InequalityConstraint if self._ineq else ConstraintFlag.EqualityConstraint

    def auto_span(self):  # src/constraints.cpp:38-43,99-104,163-169
        md = max([self._f.shape[0]] + [m.shape[0] for m in (self._E, self._G) if m is not None])
        if self._E is not None:
            self._E = AutoSpan.span_matrix(self._E, md, 1 if self._kind == "mixed" else 0)
        if self._G is not None:
            self._G = AutoSpan.span_matrix(self._G, md)
        self._f = AutoSpan.span_vector(self._f, md)


class TrajectoryConstraint(_EqIneq):
    _kind = "trajectory"

    def __init__(self, E, f, is_inequality_constraint=True):
        super().__init__("Trajectory", is_inequality_constraint)
        self._E = np.atleast_2d(np.asarray(E, dtype=np.float64)).copy()
        self._f = np.asarray(f, dtype=np.float64).reshape(-1).copy()


class ControlConstraint(_EqIneq):
    _kind = "control"

    def __init__(self, G, f, is_inequality_constraint=True):
        super().__init__("Control", is_inequality_constraint)
        self._G = np.atleast_2d(np.asarray(G, dtype=np.float64)).copy()
        self._f = np.asarray(f, dtype=np.float64).reshape(-1).copy()


class MixedConstraint(_EqIneq):
    _kind = "mixed"

    def __init__(self, E, G, f, is_inequality_constraint=True):
        super().__init__("Mixed", is_inequality_constraint)
        self._E = np.atleast_2d(np.asarray(E, dtype=np.float64)).copy()
        self._G = np.atleast_2d(np.asarray(G, dtype=np.float64)).copy()
        self._f = np.asarray(f, dtype=np.float64).reshape(-1).copy()


class TrajectoryBoundConstraint(Constraint):
    _kind = "trajectory_bound"

    def __init__(self, lower, upper):
        super().__init__("Trajectory bound constraint")
        self._lower = np.asarray(lower, dtype=np.float64).reshape(-1).copy()
        self._upper = np.asarray(upper, dtype=np.float64).reshape(-1).copy()
        if self._lower.shape != self._upper.shape:
            _domain_error("lower and upper should have the same number of rows")

    def constraint_type(self):
        return ConstraintFlag.InequalityConstraint

    def auto_span(self):
        md = max(self._lower.shape[0], self._upper.shape[0])
        self._lower, self._upper = AutoSpan.span_vector(self._lower, md), AutoSpan.span_vector(self._upper, md)


class ControlBoundConstraint(Constraint):
    _kind = "control_bound"

    def __init__(self, lower, upper):
        super().__init__("Control bound constraint")
        self._lower = np.asarray(lower, dtype=np.float64).reshape(-1).copy()
        self._upper = np.asarray(upper, dtype=np.float64).reshape(-1).copy()
        if self._lower.shape != self._upper.shape:
            _domain_error("lower and upper should have the same number of rows")

    def constraint_type(self):
        return ConstraintFlag.BoundConstraint

    def auto_span(self):
        md = max(self._lower.shape[0], self._upper.shape[0])
        self._lower, self._upper = AutoSpan.span_vector(self._lower, md), AutoSpan.span_vector(self._upper, md)

    def lower(self):
        return self._lower

    def upper(self):
        return self._upper


# ------------------------------------------------------------------------------------------- controllers
class LMPC:
    """include/LMPC.h:34-185 / src/LMPC.cpp; solve() = copra_b200_lmpc_run on a batch of one"""

    _initial_state = False

    def __init__(self, *args):
        self._ps = None
        self._costs, self._cstrs = [], []
        self._control = self._trajectory = None
        self._solve_time = self._solve_and_build_time = 0.0
        self._status, self._iters = -1, (0, 0)
        flag = SolverFlag.DEFAULT
        for a in args:
            if isinstance(a, PreviewSystem):
                self.initialize_controller(a)
            elif isinstance(a, SolverFlag):
                flag = a
            else:
                raise TypeError("LMPC([PreviewSystem], [SolverFlag])")
        self.select_qp_Solver(flag)

    def select_qp_Solver(self, flag):
        if not isinstance(flag, SolverFlag):
            raise TypeError("expected a SolverFlag")
        self._flag = flag

    def initialize_controller(self, ps):
        self._ps = ps
        self._costs, self._cstrs = [], []
        self._control = np.zeros(ps.full_u_dim)
        self._trajectory = np.zeros(ps.full_x_dim)

    def add_cost(self, cost):
        cost.initialize_cost(self._ps)  # raises RuntimeError on a dimension mismatch, like addCost (src/LMPC.cpp:118-122)
        self._costs.append(cost)

    def add_constraint(self, cstr):
        cstr.initialize_constraint(self._ps)
        self._cstrs.append(cstr)

    def reset_constraints(self):
        self._cstrs = []

    def _problem(self):
        ps = self._ps
        return dict(ps._base(), initial_state=self._initial_state, costs=[c._desc() for c in self._costs],
                    constraints=[c._desc() for c in self._cstrs])

    def solve(self):
        t0 = time.perf_counter()
        eng = _eng()
        out = eng.lmpc_run(self._problem(), want=("control", "trajectory", "status", "iters"))
        self._status = int(out["status"][0])
        self._iters = tuple(int(v) for v in out["iters"][0])
        ok = self._status == 0
        if ok:
            self._control, self._trajectory = out["control"][0].copy(), out["trajectory"][0].copy()
        self._ps.is_updated = True
        self._solve_time = eng.timing()["solve_ms"] * 1e-3
        self._purge()
        self._solve_and_build_time = time.perf_counter() - t0
        return ok

    def _purge(self):
        """costs / constraints only the controller still refers to are removed AFTER the solve, with the reference's
        warning (src/LMPC.cpp:288-307: shared_ptr::use_count() == 1)"""
        for lst in (self._costs, self._cstrs):
            i = 0
            while i < len(lst):
                # references: the list + the getrefcount argument
                if sys.getrefcount(lst[i]) <= 2:
                    sys.stderr.write("A '%s' has been destroyed.\nIt has been removed from the controller\n" % lst[i].name())
                    del lst[i]
                else:
                    i += 1

    def control(self):
        return self._control

    def trajectory(self):
        return self._trajectory

    def solve_time(self):
        return self._solve_time

    def solve_and_build_time(self):
        return self._solve_and_build_time

    def fail_code(self):  # SolverInterface::SI_fail
        return self._status

    def iterations(self):
        return self._iters


class InitialStateLMPC(LMPC):
    """include/InitialStateLMPC.h / src/InitialStateLMPC.cpp: the initial state joins the decision variables"""

    _initial_state = True

    def __init__(self, *args):
        self._R = self._r = self._x0lb = self._x0ub = None
        self._x0_result = None
        super().__init__(*args)

    def initialize_controller(self, ps):
        super().initialize_controller(ps)
        nx = ps.x_dim
        self._R, self._r = np.zeros((nx, nx)), np.zeros(nx)
        # reference defaults (src/InitialStateLMPC.cpp:21-28): both bounds are ps->x0, i.e. x0 stays pinned until
        # reset_initial_state_bounds is called (None = let the C ABI apply that default from the current x0)
        self._x0lb = self._x0ub = None

    def reset_initial_state_cost(self, R, r):
        self._R = np.atleast_2d(np.asarray(R, dtype=np.float64)).copy()
        self._r = np.asarray(r, dtype=np.float64).reshape(-1).copy()

    def reset_initial_state_bounds(self, lower, upper):
        self._x0lb = np.asarray(lower, dtype=np.float64).reshape(-1).copy()
        self._x0ub = np.asarray(upper, dtype=np.float64).reshape(-1).copy()

    def _problem(self):
        return dict(super()._problem(), R=self._R, r=self._r, x0lb=self._x0lb, x0ub=self._x0ub)

    def solve(self):
        t0 = time.perf_counter()
        eng = _eng()
        out = eng.lmpc_run(self._problem(), want=("control", "trajectory", "x", "status", "iters"))
        self._status = int(out["status"][0])
        self._iters = tuple(int(v) for v in out["iters"][0])
        ok = self._status == 0
        if ok:
            self._control, self._trajectory = out["control"][0].copy(), out["trajectory"][0].copy()
            self._x0_result = out["x"][0][:self._ps.x_dim].copy()
        self._ps.is_updated = True
        self._solve_time = eng.timing()["solve_ms"] * 1e-3
        self._purge()
        self._solve_and_build_time = time.perf_counter() - t0
        return ok

    def initial_state(self):
        return self._x0_result
