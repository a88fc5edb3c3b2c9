"""The reference's python test scenarios (binding/python/tests/pyTests.py) against copra_b200.pycopra, which mirrors the
pyCopra module surface (binding/python/CopraBindings.cpp).  The scenario data are the reference's BoundedSystem /
IneqSystem / EqSystem fixtures (tests/systems.h); results are additionally checked against the CPU oracle."""
import gc

import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def copra():
    from copra_b200 import capi, pycopra
    if capi.load().copra_b200_device_count() < 1:
        pytest.skip("no CUDA device")
    return pycopra


class Fx:  # pyTests.py setUp
    T, mass, N = 0.005, 5.0, 300
    A = np.array([[1.0, T], [0.0, 1.0]])
    B = np.array([[0.5 * T * T / mass], [T / mass]])
    c = np.array([(-9.81 / 2.0) * T ** 2, -9.81 * T])
    x0 = np.array([0.0, -5.0])
    wu, wx = np.array([1e-4]), np.array([10.0, 10000.0])
    xd, ud = np.zeros(2), np.zeros(1)
    M, Nm = np.identity(2), np.ones((1, 1))
    Gineq, hineq = np.ones((1, 1)), np.array([200.0])
    Eineq, fineq = np.array([[0.0, 1.0]]), np.zeros(1)
    uLower, uUpper = np.array([-np.inf]), np.array([200.0])
    xLower, xUpper = np.array([-np.inf, -np.inf]), np.array([np.inf, 0.0])
    Geq, heq = np.ones((1, 1)), np.array([200.0])
    Eeq, feq = np.array([[1.0, 0.0], [0.0, 0.0]]), np.zeros(2)


def _controller(copra, costs=True):
    ps = copra.PreviewSystem()
    ps.system(Fx.A, Fx.B, Fx.c, Fx.x0, Fx.N)
    ctl = copra.LMPC(ps)
    x_cost, u_cost = copra.TargetCost(Fx.M, -Fx.xd), copra.ControlCost(Fx.Nm, -Fx.ud)
    x_cost.weights(Fx.wx)
    u_cost.weights(Fx.wu)
    if costs:
        ctl.add_cost(x_cost)
        ctl.add_cost(u_cost)
    return ps, ctl, x_cost, u_cost


def _split(traj):
    return traj[0::2], traj[1::2]


def _oracle(constraints):
    prob = dict(nx=2, nu=1, N=Fx.N, A=Fx.A, B=Fx.B, d=Fx.c, x0=Fx.x0, initial_state=False,
                costs=[dict(kind="target", M=Fx.M, p=-Fx.xd, w=Fx.wx), dict(kind="control", N=Fx.Nm, p=-Fx.ud, w=Fx.wu)],
                constraints=constraints)
    return po.lmpc(prob)


def test_lmpc_ineq(copra):
    ps, ctl, xc, uc = _controller(copra)
    traj_c, cont_c = copra.TrajectoryConstraint(Fx.Eineq, Fx.fineq), copra.ControlConstraint(Fx.Gineq, Fx.hineq)
    ctl.add_constraint(traj_c)
    ctl.add_constraint(cont_c)
    assert ctl.solve()
    pos, vel = _split(ctl.trajectory())
    assert abs(Fx.xd[1] - vel[-1]) < 5e-4
    assert pos.max() <= Fx.x0[0] and ctl.control().max() <= Fx.hineq[0] + 1e-9
    assert ctl.solve_time() > 0 and ctl.solve_and_build_time() >= ctl.solve_time()
    o = _oracle([dict(kind="trajectory", E=Fx.Eineq, f=Fx.fineq), dict(kind="control", G=Fx.Gineq, f=Fx.hineq)])
    assert np.abs(ctl.control() - o["control"]).max() < 1e-6 * max(1.0, np.abs(o["control"]).max())
    assert ctl.iterations() == tuple(o["iter"])


def test_lmpc_mixed(copra):
    ps, ctl, xc, uc = _controller(copra)
    mixed = copra.MixedConstraint(Fx.Eineq, Fx.Gineq, Fx.hineq)
    ctl.add_constraint(mixed)
    assert ctl.solve()
    traj, u = ctl.trajectory(), ctl.control()
    pos, vel = _split(traj)
    assert abs(Fx.xd[1] - vel[-1]) < 5e-4 and pos.max() <= Fx.x0[0]
    for i in range(Fx.N):
        assert Fx.Eineq[0] @ traj[2 * i:2 * i + 2] + Fx.Gineq[0, 0] * u[i] <= Fx.hineq[0] + 1e-6


def test_lmpc_bound(copra):
    ps, ctl, xc, uc = _controller(copra)
    tb, cb = copra.TrajectoryBoundConstraint(Fx.xLower, Fx.xUpper), copra.ControlBoundConstraint(Fx.uLower, Fx.uUpper)
    ctl.add_constraint(tb)
    ctl.add_constraint(cb)
    assert ctl.solve()
    pos, vel = _split(ctl.trajectory())
    assert abs(Fx.xd[1] - vel[-1]) < 5e-4 and pos.max() <= Fx.x0[0]
    assert vel.max() <= Fx.xUpper[1] + 1e-6 and ctl.control().max() <= Fx.uUpper[0] + 1e-6
    o = _oracle([dict(kind="trajectory_bound", lower=Fx.xLower, upper=Fx.xUpper), dict(kind="control_bound", lower=Fx.uLower, upper=Fx.uUpper)])
    assert ctl.iterations() == tuple(o["iter"])
    assert np.abs(ctl.control() - o["control"]).max() < 1e-6 * max(1.0, np.abs(o["control"]).max())


def test_lmpc_eq(copra):
    ps = copra.PreviewSystem()
    x0 = np.array([0.0, 0.0])
    ps.system(Fx.A, Fx.B, Fx.c, x0, Fx.N)
    ctl = copra.LMPC(ps)
    xc, uc = copra.TargetCost(Fx.M, -Fx.xd), copra.ControlCost(Fx.Nm, -Fx.ud)
    xc.weights(Fx.wx)
    uc.weights(Fx.wu)
    eq = copra.TrajectoryConstraint(Fx.Eeq, Fx.feq, False)
    ctl.add_cost(xc)
    ctl.add_cost(uc)
    ctl.add_constraint(eq)
    assert ctl.solve()
    pos, vel = _split(ctl.trajectory())
    assert abs(vel[-1]) < 5e-4 and np.abs(pos).max() <= 1e-6  # position pinned by the equality rows


def test_constructors_and_failures(copra):
    ps = copra.PreviewSystem()
    ps.system(Fx.A, Fx.B, Fx.c, Fx.x0, Fx.N)
    ctl = copra.LMPC(ps)
    copra.LMPC()
    copra.LMPC(copra.SolverFlag.QuadProgDense)
    copra.LMPC(ps, copra.SolverFlag.QuadProgDense)
    ctl.initialize_controller(ps)
    for cls in (copra.TrajectoryConstraint, copra.ControlConstraint, copra.TrajectoryBoundConstraint, copra.ControlBoundConstraint):
        with pytest.raises(TypeError):
            cls()
    with pytest.raises(RuntimeError):
        copra.PreviewSystem(Fx.A, Fx.B, Fx.c, Fx.x0, 0)


def test_throw_handler(copra):
    ps, ctl, xc, uc = _controller(copra, costs=False)
    bad = [copra.TrajectoryConstraint(np.identity(5), np.ones(2)), copra.ControlConstraint(np.identity(5), np.ones(2)),
           copra.MixedConstraint(np.identity(5), np.identity(5), np.ones(2)),
           copra.TrajectoryBoundConstraint(np.ones(3), np.ones(3)), copra.ControlBoundConstraint(np.ones(3), np.ones(3))]
    for c in bad:
        with pytest.raises(RuntimeError):
            ctl.add_constraint(c)
    with pytest.raises(RuntimeError):
        copra.TrajectoryBoundConstraint(np.ones(3), np.ones(2))
    with pytest.raises(RuntimeError):
        ctl.add_cost(copra.TargetCost(np.identity(5), np.ones(2)))


def test_constraint_and_cost_deletion(copra, capfd):
    """costs / constraints the caller dropped are removed after the next solve (use_count rule, src/LMPC.cpp:288-307)"""
    ps = copra.PreviewSystem()
    ps.system(Fx.A, Fx.B, Fx.c, Fx.x0, Fx.N)
    ctl = copra.LMPC(ps)
    traj_c = copra.TrajectoryConstraint(Fx.Eineq, Fx.fineq)
    cont_c = copra.ControlConstraint(Fx.Gineq, Fx.hineq)
    traj_eq = copra.TrajectoryConstraint(Fx.Eeq, Fx.feq, False)
    cont_eq = copra.ControlConstraint(Fx.Geq, Fx.heq, False)
    traj_bd = copra.TrajectoryBoundConstraint(Fx.xLower, Fx.xUpper)
    cont_bd = copra.ControlBoundConstraint(Fx.uLower, Fx.uUpper)
    target, trajectory = copra.TargetCost(Fx.M, -Fx.xd), copra.TrajectoryCost(Fx.M, -Fx.xd)
    control, mixed = copra.ControlCost(Fx.Nm, -Fx.ud), copra.MixedCost(np.ones((1, 2)), Fx.Nm, -Fx.ud)
    for c in (traj_c, cont_c, traj_eq, cont_eq, traj_bd, cont_bd):
        ctl.add_constraint(c)
    for c in (target, trajectory, control, mixed):
        ctl.add_cost(c)
    del traj_c
    target.weights(Fx.wx)
    control.weights(Fx.wu)
    del traj_eq, cont_eq, traj_bd, cont_bd, trajectory, mixed, c
    gc.collect()
    assert not ctl.solve()  # u == 200 at every step contradicts the trajectory constraints: infeasible
    err = capfd.readouterr().err
    assert err.count("has been destroyed") == 7
    assert ctl.solve()      # only cont_c, target and control are left


def test_preview_system_outlives_its_name(copra):
    ps, ctl, xc, uc = _controller(copra)
    del ps
    gc.collect()
    a, b = copra.TrajectoryConstraint(Fx.Eineq, Fx.fineq), copra.ControlConstraint(Fx.Gineq, Fx.hineq)
    ctl.add_constraint(a)
    ctl.add_constraint(b)
    assert ctl.solve()
    assert ctl.control().max() <= Fx.hineq[0] + 1e-9


def test_autospan_full_size_entries(copra):
    """time-varying (full-size) control bound rows through auto_span: same result as the step-size entry"""
    ps, ctl, xc, uc = _controller(copra)
    G, h = copra.AutoSpan.span_matrix(Fx.Gineq, Fx.N), copra.AutoSpan.span_vector(Fx.hineq, Fx.N)
    assert G.shape == (Fx.N, Fx.N) and h.shape == (Fx.N,)
    full = copra.ControlConstraint(G, h)
    traj_c = copra.TrajectoryConstraint(Fx.Eineq, Fx.fineq)
    traj_c.auto_span()  # nothing to do: E and f agree
    ctl.add_constraint(full)
    ctl.add_constraint(traj_c)
    assert ctl.solve()
    u_full = ctl.control().copy()
    ps2, ctl2, xc2, uc2 = _controller(copra)
    s1, s2 = copra.TrajectoryConstraint(Fx.Eineq, Fx.fineq), copra.ControlConstraint(Fx.Gineq, Fx.hineq)
    ctl2.add_constraint(s2)
    ctl2.add_constraint(s1)
    assert ctl2.solve()
    assert np.abs(u_full - ctl2.control()).max() < 1e-6 * max(1.0, np.abs(u_full).max())


def test_single_entry_getters_and_initial_state(copra):
    ps, ctl, xc, uc = _controller(copra)
    ps.update_system()
    assert ps.is_updated and ps.Psi.shape == (2 * (Fx.N + 1), Fx.N)
    uc.update(ps)
    assert np.allclose(uc.Q(), np.diag(np.full(Fx.N, Fx.wu[0])))  # N'WN on the block diagonal, no regulariser
    tc = copra.TrajectoryConstraint(Fx.Eineq, Fx.fineq)
    tc.initialize_constraint(ps)
    tc.update(ps)
    assert tc.nr_constr() == Fx.N + 1 and tc.A().shape == (Fx.N + 1, Fx.N)
    assert np.allclose(tc.A(), ps.Psi[1::2, :])
    isl = copra.InitialStateLMPC(ps)
    a, b = copra.TargetCost(Fx.M, -Fx.xd), copra.ControlCost(Fx.Nm, -Fx.ud)
    a.weights(np.array([10.0, 100.0]))
    b.weights(np.array([1e-2]))
    cb = copra.ControlBoundConstraint(Fx.uLower, Fx.uUpper)
    isl.add_cost(a)
    isl.add_cost(b)
    isl.add_constraint(cb)
    isl.reset_initial_state_cost(np.identity(2), -Fx.x0)
    isl.reset_initial_state_bounds(Fx.x0 - 0.1, Fx.x0 + 0.1)
    assert isl.solve()
    assert np.all(np.abs(isl.initial_state() - Fx.x0) <= 0.1 + 1e-9)


def test_initial_state_default_bounds_pin_x0(copra):
    """reference constructor (src/InitialStateLMPC.cpp:21-28): x0lb = x0ub = ps->x0 until resetInitialStateBounds is called,
    so with only an initial-state cost the optimal x0 is the given one"""
    ps, _, _, _ = _controller(copra)
    isl = copra.InitialStateLMPC(ps)
    a, b = copra.TargetCost(Fx.M, -Fx.xd), copra.ControlCost(Fx.Nm, -Fx.ud)
    a.weights(np.array([10.0, 100.0]))
    b.weights(np.array([1e-2]))
    isl.add_cost(a)
    isl.add_cost(b)
    isl.add_constraint(copra.ControlBoundConstraint(Fx.uLower, Fx.uUpper))
    isl.reset_initial_state_cost(np.identity(2), np.array([3.0, -7.0]))  # pulls x0 away; the default bounds hold it
    assert isl.solve()
    assert np.abs(isl.initial_state() - Fx.x0).max() <= 1e-9
