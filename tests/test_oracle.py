"""CPU tests (-m "not gpu"): the oracle against golden vectors / known answers / the independent numpy
restatement, the host-side logic, and that the C-ABI library loads and exports every declared symbol."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

from copra_b200 import workloads as wl
from oracle import copra_numpy as cn
from oracle import pyoracle as po
from tests.util import rel_err, x_err

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STAGES = ("Phi", "Psi", "xi", "Q", "c", "Aeq", "beq", "Aineq", "bineq", "lb", "ub")


def _ka():
    return json.load(open(os.path.join(HERE, "golden", "ka_problem.json")))


def test_ka1_known_answer():
    """reference fixture `Problem` (tests/systems.h:11-38): x*, objective, active set, iterations"""
    ka = _ka()
    r = po.quadprog(ka["Q"], ka["c"], ka["Aeq"], ka["beq"], ka["Aineq"], ka["bineq"], ka["XL"], ka["XU"])
    assert r["fail"] == 0
    assert np.abs(r["x"] - np.array(ka["x"])).max() < 5e-15
    assert abs(r["crval"] - ka["objective"]) < 1e-12
    assert list(r["iact"]) == ka["iact"] and list(r["iter"]) == ka["iter"]
    # multipliers (quadprog convention: positive for active rows)
    lag = r["lagr"]
    assert np.allclose(np.abs(lag[:3]), np.abs(ka["multipliers"]["eq"]), atol=1e-12)
    assert abs(lag[3] - ka["multipliers"]["ineq0"]) < 1e-12


def test_c1_reference_points():
    """iteration counts / active rows / objective of the BoundedSystem fixture"""
    pts = json.load(open(os.path.join(HERE, "golden", "c1_reference_points.json")))
    assert pts["target"]["iter"] == [113, 0] and pts["target"]["nact"] == 112
    assert pts["trajectory"]["iter"] == [27, 0] and pts["trajectory"]["nact"] == 26
    o = po.lmpc(wl.instance(wl.c1("target"), 0))
    assert list(o["iter"]) == pts["target"]["iter"] and o["nact"] == pts["target"]["nact"]
    assert abs(o["crval"] - (-1752900.49379)) < 1e-4
    assert abs(o["trajectory"][-1] - (-1.000129)) < 1e-6
    # the properties the reference test asserts (tests/TestLMPC.cpp:78-83)
    traj, u = o["trajectory"], o["control"]
    assert abs(-1.0 - traj[-1]) <= 1e-3 and traj[0::2].max() <= 0.0 and traj[1::2].max() <= 1e-6 and u.max() <= 200 + 1e-6


def _golden_cases():
    z = np.load(os.path.join(HERE, "golden", "small_cases.npz"))
    probs = json.load(open(os.path.join(HERE, "golden", "small_cases_problems.json")))
    for name, prob in probs.items():
        yield name, _revive(prob), {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(name + "/")}


def _revive(prob):
    def arr(v):
        return None if v is None else np.array([[np.nan if e is None else e for e in row] if isinstance(row, list) else row
                                                for row in v], dtype=float) if isinstance(v, list) else v
    out = dict(prob)
    for k in ("A", "B", "d", "x0", "R", "r", "x0lb", "x0ub"):
        if out.get(k) is not None:
            out[k] = np.asarray(out[k], dtype=float)
    out["costs"] = [{k: (np.asarray(v, dtype=float) if isinstance(v, list) else v) for k, v in c.items()} for c in prob["costs"]]
    out["constraints"] = [{k: (np.asarray(v, dtype=float) if isinstance(v, list) else v) for k, v in c.items()}
                          for c in prob["constraints"]]
    return out


@pytest.mark.parametrize("name,prob,gold", list(_golden_cases()), ids=[n for n, _, _ in _golden_cases()])
def test_oracle_matches_committed_goldens(name, prob, gold):
    o = po.lmpc(prob)
    for k in STAGES + ("x", "control", "trajectory"):
        assert rel_err(o[k], gold[k]) <= 1e-14, (name, k)
    assert list(o["iact"]) == list(gold["iact"]) and list(o["iter"]) == list(gold["iter"])


@pytest.mark.parametrize("name,prob,gold", list(_golden_cases()), ids=[n for n, _, _ in _golden_cases()])
def test_numpy_restatement_agrees(name, prob, gold):
    """independent closed-form restatement of K1-K5 vs the C++ oracle's goldens + KKT of the golden solution"""
    q = cn.build_qp(prob)
    tol = dict(Q=1e-9) if prob.get("initial_state") else {}  # LU-inverse vs numpy inverse of the Schur block (quirk Q7)
    for k in STAGES:
        assert rel_err(q[k], gold[k]) <= tol.get(k, 1e-12), (name, k, rel_err(q[k], gold[k]))
    kkt = cn.kkt_residuals(gold["Q"], gold["c"], gold["Aeq"], gold["beq"], gold["Aineq"], gold["bineq"], gold["lb"], gold["ub"],
                           gold["x"], gold["iact"])
    assert kkt["stationarity"] <= 1e-8 and kkt["primal"] <= 1e-9 and kkt["dual"] <= 1e-9 and kkt["complementarity"] <= 1e-9, kkt


def test_oracle_vs_numpy_on_configs():
    for bp in (wl.c2(batch=3), wl.c3(batch=2, N=30), wl.c4(batch=2), wl.c5(batch=1, N=20)):
        for i in range(bp["batch"]):
            prob = wl.instance(bp, i)
            o, q = po.lmpc(prob), cn.build_qp(prob)
            for k in STAGES:
                assert rel_err(q[k], o[k]) <= (1e-9 if (k == "Q" and bp["initial_state"]) else 1e-12), (bp["name"], k)
            kkt = cn.kkt_residuals(o["Q"], o["c"], o["Aeq"], o["beq"], o["Aineq"], o["bineq"], o["lb"], o["ub"], o["x"], o["iact"])
            assert max(kkt["stationarity"], kkt["primal"], kkt["dual"], kkt["complementarity"]) <= 1e-8, (bp["name"], kkt)


def test_random_qps_kkt_and_drops():
    """the GI restatement on random QPs with equalities and mixed finite / inf / DBL_MAX bounds"""
    rng = np.random.default_rng(3)
    drops = 0
    for trial in range(60):
        n = int(rng.integers(2, 20))
        meq, m = int(rng.integers(0, max(1, n // 3))), int(rng.integers(0, 2 * n))
        L = rng.normal(size=(n, n))
        Q, c = L @ L.T + 0.1 * np.eye(n), rng.normal(size=n)
        xf = rng.normal(size=n)
        Aeq, Aineq = rng.normal(size=(meq, n)), rng.normal(size=(m, n))
        beq, bineq = Aeq @ xf, Aineq @ xf + rng.uniform(0, 1, m)
        lb, ub = xf - rng.uniform(0.1, 2, n), xf + rng.uniform(0.1, 2, n)
        lb[rng.uniform(size=n) < 0.3] = -np.inf
        ub[rng.uniform(size=n) < 0.3] = np.finfo(float).max
        r = po.quadprog(Q, c, Aeq, beq, Aineq, bineq, lb, ub)
        assert r["fail"] == 0
        kkt = cn.kkt_residuals(Q, c, Aeq, beq, Aineq, bineq, lb, ub, r["x"], r["iact"])
        assert max(kkt["stationarity"], kkt["primal"], kkt["dual"], kkt["complementarity"]) <= 1e-9, (trial, kkt)
        drops += r["iter"][1]
    assert drops > 0


def _random_qp(rng, nmax=24, tight=False):
    n = int(rng.integers(2, nmax + 1))
    meq, m = int(rng.integers(0, max(1, n // 3))), int(rng.integers(0, 2 * n))
    L = rng.normal(size=(n, n))
    Q, c = L @ L.T + 0.1 * np.eye(n), rng.normal(size=n) * (5.0 if tight else 1.0)
    xf = rng.normal(size=n)
    Aeq, Aineq = rng.normal(size=(meq, n)), rng.normal(size=(m, n))
    beq, bineq = Aeq @ xf, Aineq @ xf + rng.uniform(0, 0.2 if tight else 1.0, m)
    lb, ub = xf - rng.uniform(0.05 if tight else 0.1, 2, n), xf + rng.uniform(0.05 if tight else 0.1, 2, n)
    lb[rng.uniform(size=n) < 0.3] = -np.inf
    ub[rng.uniform(size=n) < 0.3] = np.finfo(float).max
    lb[rng.uniform(size=n) < 0.1] = -np.finfo(float).max
    ub[rng.uniform(size=n) < 0.1] = np.inf
    return Q, c, Aeq, beq, Aineq, bineq, lb, ub


def test_two_independent_k6_restatements_agree_step_for_step():
    """the C++ oracle's qpgen2 was single-sourced.  oracle/qpgen2_numpy.py restates the same published
    algorithm a second time (numpy, dense R, whole-column rotations); on 600 random QPs with equalities, mixed
    finite / inf / DBL_MAX bounds and ~1000 constraint drops both give the same x, the same active set IN ADD ORDER,
    the same outer-iteration and drop counts and the same fail code."""
    from oracle import qpgen2_numpy as q2
    rng = np.random.default_rng(20261017)
    drops = outer = 0
    for trial in range(600):
        args = _random_qp(rng, tight=trial % 2 == 1)
        a = po.quadprog(*args)
        b = q2.solve_copra_qp(*args)
        assert a["fail"] == b["fail"], trial
        if a["fail"] != 0:
            continue
        assert np.abs(a["x"] - b["x"]).max() <= 1e-9 * max(1.0, np.abs(a["x"]).max()), trial
        assert [int(i) for i in a["iact"]] == b["iact"], (trial, list(a["iact"]), b["iact"])
        assert tuple(int(v) for v in a["iter"]) == tuple(b["iter"]), (trial, a["iter"], b["iter"])
        assert abs(a["crval"] - b["crval"]) <= 1e-9 * max(1.0, abs(a["crval"])), trial
        drops += b["iter"][1]
        outer += b["iter"][0]
    assert drops >= 900 and outer >= 6000, (drops, outer)


def test_k6_restatements_agree_on_lmpc_configs():
    """same cross-check on assembled LMPC QPs: KA-1, the C2 shape (35 outer iterations, trajectory-bound rows active), the
    C4 shape (InitialStateLMPC: ~50 drops) and an infeasible / a non-PD problem"""
    from oracle import qpgen2_numpy as q2
    ka = _ka()
    b = q2.solve_copra_qp(ka["Q"], ka["c"], ka["Aeq"], ka["beq"], ka["Aineq"], ka["bineq"], ka["XL"], ka["XU"])
    assert b["fail"] == 0 and np.abs(b["x"] - np.array(ka["x"])).max() < 1e-13
    assert b["iact"] == ka["iact"] and list(b["iter"]) == ka["iter"] and abs(b["crval"] - ka["objective"]) < 1e-12
    for bp in (wl.c2(batch=3), wl.c4(batch=2)):
        for i in range(bp["batch"]):
            o = po.lmpc(wl.instance(bp, i))
            r = q2.solve_copra_qp(o["Q"], o["c"], o["Aeq"], o["beq"], o["Aineq"], o["bineq"], o["lb"], o["ub"])
            assert r["fail"] == o["fail"] == 0
            assert x_err(r["x"], o["x"]) <= 1e-8, (bp["name"], i, x_err(r["x"], o["x"]))
            assert sorted(r["iact"]) == sorted(int(k) for k in o["iact"]), (bp["name"], i)
            assert r["iter"][0] == int(o["iter"][0]) and abs(r["iter"][1] - int(o["iter"][1])) <= 2, (bp["name"], i, r["iter"], o["iter"])
    r = q2.solve_copra_qp(np.eye(2), np.zeros(2), None, None, [[1.0, 0.0], [-1.0, 0.0]], [-1.0, -1.0], [-np.inf] * 2, [np.inf] * 2)
    assert r["fail"] == 1
    r = q2.solve_copra_qp(np.array([[1.0, 2.0], [2.0, 1.0]]), np.zeros(2), None, None, None, None, [-1.0] * 2, [1.0] * 2)
    assert r["fail"] == 2


def test_oracle_fail_codes_and_exceptions():
    r = po.quadprog(np.eye(2), np.zeros(2), None, None, [[1.0, 0.0], [-1.0, 0.0]], [-1.0, -1.0], [-np.inf] * 2, [np.inf] * 2)
    assert r["fail"] == 1
    r = po.quadprog([[1.0, 2.0], [2.0, 1.0]], np.zeros(2), None, None, None, None, [-1.0] * 2, [1.0] * 2)
    assert r["fail"] == 2
    bp = wl.c2(batch=1)
    bad = wl.instance(bp, 0)
    bad["costs"][0]["M"] = np.eye(3)
    with pytest.raises(po.OracleError) as e:
        po.lmpc(bad)
    assert e.value.code == -1  # std::domain_error
    bad = wl.instance(bp, 0)
    bad["N"] = 0
    with pytest.raises(po.OracleError):
        po.lmpc(bad)


def test_reference_quirks_are_restated():
    """Q1 (lower trajectory bounds are not negated), Q4 (DBL_MAX default bounds), Q5 (step-0 rows), Q6 (mixed: N steps)"""
    bp = wl.c2(batch=1, N=6)
    prob = wl.instance(bp, 0)
    prob["constraints"] = [dict(kind="trajectory_bound", lower=np.array([-np.inf, -3.0]), upper=np.array([np.inf, 0.0]))]
    o = po.lmpc(prob, solve=False)
    assert o["Aineq"].shape[0] == 2 * 7
    assert np.array_equal(o["Aineq"][:7], o["Aineq"][7:])          # Q1: same +Psi rows for lower and upper lines
    assert np.all(o["Aineq"][0] == 0.0) and np.all(o["Aineq"][7] == 0.0)  # Q5: structurally zero step-0 rows
    assert np.all(o["lb"] == -np.finfo(float).max) and np.all(o["ub"] == np.finfo(float).max)  # Q4
    prob["constraints"] = [dict(kind="mixed", E=np.array([[0.0, 1.0]]), G=np.array([[1.0]]), f=np.array([5.0]))]
    assert po.lmpc(prob, solve=False)["Aineq"].shape[0] == 6       # Q6


def test_workloads_are_deterministic_and_shaped():
    a, b = wl.c3(batch=5), wl.c3(batch=5)
    assert np.array_equal(a["x0"], b["x0"]) and np.array_equal(a["constraints"][0]["E"], b["constraints"][0]["E"])
    assert not np.array_equal(wl.c3(batch=5, seed_offset=1)["x0"], a["x0"])
    sizes = {"c1": (300, 301), "c2": (50, 51), "c3": (320, 640), "c4": (52, 0), "c5": (800, 1608)}
    for name, (n, m) in sizes.items():
        bp = wl.CONFIGS[name]() if name == "c1" else wl.CONFIGS[name](batch=1)
        s = po.sizes(wl.instance(bp, 0))
        assert (s["nvar"], s["mineq"]) == (n, m), (name, s)
    # SplitMix64 reference value (seed 0 first output)
    assert int(wl.SplitMix64(0).u64(1)[0]) == 0xE220A8397B1DCDAF
    sub, (lo, hi) = wl.shard(wl.c2(batch=10), 1, 4)
    assert (lo, hi) == (3, 6) and sub["batch"] == 3


def test_library_exports_every_declared_symbol():
    """the C-ABI library loads (no GPU needed) and exports exactly what include/copra_b200.h declares"""
    from copra_b200 import capi
    lib = capi.load()
    header = open(os.path.join(ROOT, "include", "copra_b200.h")).read()
    declared = set(re.findall(r"\b(copra_b200_[a-z_0-9]+)\s*\(", header))
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.copra_b200_abi_version() == 1
    out = subprocess.run(["nm", "-D", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in out, "the product library must not contain or link the oracle"
    if lib.copra_b200_device_count() == 0:  # no CPU fallback: creation must fail loudly
        with pytest.raises(capi.CopraB200Error):
            capi.Engine(0)


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "copra_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "copra_oracle" not in txt, os.path.join(dirpath, f)
    for f in os.listdir(os.path.join(ROOT, "include", "copra")):
        assert "oracle" not in open(os.path.join(ROOT, "include", "copra", f)).read()


def test_facade_host_logic():
    """reference error-handling and autoSpan scenarios against the C++ facade (no GPU needed)"""
    exe = os.path.join(ROOT, "tests", "cpp", "test_facade")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", ROOT, "-s", "tests/cpp/test_facade"])
    r = subprocess.run([exe, "cpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failures" in r.stdout


def test_pycopra_autospan_and_system_checks_cpu():
    """host-side pieces of the pyCopra-compatible front end need no GPU: AutoSpan (src/AutoSpan.cpp:10-48) and the
    PreviewSystem dimension checks (src/PreviewSystem.cpp:27-49)"""
    from copra_b200 import pycopra as copra
    m = copra.AutoSpan.span_matrix(np.array([[1.0, 2.0]]), 3)
    assert m.shape == (3, 6) and np.array_equal(m[1], [0, 0, 1, 2, 0, 0])
    assert copra.AutoSpan.span_matrix(np.array([[1.0, 2.0]]), 2, 1).shape == (2, 6)  # one extra zero block column
    assert np.array_equal(copra.AutoSpan.span_vector(np.array([1.0, 2.0]), 6), [1, 2, 1, 2, 1, 2])
    with pytest.raises(RuntimeError):
        copra.AutoSpan.span_matrix(np.ones((2, 2)), 5)
    ps = copra.PreviewSystem()
    with pytest.raises(RuntimeError):
        ps.system(np.eye(2), np.ones((2, 1)), np.zeros(2), np.zeros(2), 0)
    with pytest.raises(RuntimeError):
        ps.system(np.eye(2), np.ones((3, 1)), np.zeros(2), np.zeros(2), 4)
    ps.system(np.eye(2), np.ones((2, 1)), np.zeros(2), np.zeros(2), 4)
    assert (ps.x_dim, ps.u_dim, ps.nr_x_Step, ps.full_x_dim, ps.full_u_dim) == (2, 1, 5, 10, 4) and not ps.is_updated
    cost = copra.MixedCost(np.ones((1, 2)), np.ones((1, 1)), np.array([1.0, 2.0, 3.0]))
    cost.auto_span()
    assert cost._M.shape == (3, 8) and cost._N.shape == (3, 3) and cost._w.shape == (3,)


def test_pycopra_dimension_errors_need_no_gpu():
    """the reference's test_throw_handler scenario (binding/python/tests/pyTests.py:311-339): shape mismatches are
    rejected on the host, before the engine is touched, as RuntimeError"""
    from copra_b200 import pycopra as copra
    T, mass, N = 0.005, 5.0, 300
    A = np.array([[1.0, T], [0.0, 1.0]])
    B = np.array([[0.5 * T * T / mass], [T / mass]])
    ps = copra.PreviewSystem(A, B, np.zeros(2), np.array([0.0, -5.0]), N)
    ctl = copra.LMPC(ps)
    bad = [copra.TrajectoryConstraint(np.identity(5), np.ones(2)), copra.ControlConstraint(np.identity(5), np.ones(2)),
           copra.MixedConstraint(np.identity(5), np.identity(5), np.ones(2)),
           copra.TrajectoryBoundConstraint(np.ones(3), np.ones(3)), copra.ControlBoundConstraint(np.ones(3), np.ones(3)),
           copra.TrajectoryConstraint(np.ones((2, 3)), np.ones(2))]
    for c in bad:
        with pytest.raises(RuntimeError):
            ctl.add_constraint(c)
    for c in (copra.TargetCost(np.identity(5), np.ones(2)), copra.ControlCost(np.ones((1, 2)), np.ones(1)),
              copra.MixedCost(np.ones((1, 2)), np.ones((1, N)), np.ones(1))):
        with pytest.raises(RuntimeError):
            ctl.add_cost(c)
    with pytest.raises(RuntimeError):
        copra.TrajectoryBoundConstraint(np.ones(3), np.ones(2))
    assert ctl._costs == [] and ctl._cstrs == []
