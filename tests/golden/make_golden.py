#!/usr/bin/env python
"""Regenerates the golden fixtures under tests/golden/ from the CPU oracle (the reference itself cannot
be built offline: no Eigen / eigen-quadprog / gfortran).  The oracle is pinned by
ka_problem.json (known answer of the reference's `Problem` fixture), by the SURVEY 8c reference points
frozen in c1_reference_points.json, and by the independent numpy restatement (tests/test_oracle.py).

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from copra_b200 import workloads as wl  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ("Phi", "Psi", "xi", "Q", "c", "Aeq", "beq", "Aineq", "bineq", "lb", "ub", "x", "control", "trajectory", "iact")


def small_cases():
    """one small instance of every config shape (N shortened so the fixtures stay a few hundred kB)"""
    yield "c2_n12", wl.instance(wl.c2(batch=4, N=12, T=0.12), 1)
    yield "c3_n10", wl.instance(wl.c3(batch=4, N=10, T=0.1), 2)
    yield "c4_n12", wl.instance(wl.c4(batch=4, N=12, T=0.12), 3)
    yield "c5_n6", wl.instance(wl.c5(batch=2, N=6), 1)
    # every cost x constraint kind at once (mirrors tests/TestLMPC_InitialState.cpp's system: A = B = ones)
    nx, nu, N = 2, 1, 6
    mixed = dict(nx=nx, nu=nu, N=N, initial_state=False, A=np.array([[1.0, 0.1], [0.0, 1.0]]), B=np.array([[0.005], [1.0]]),
                 d=np.array([0.01, -0.02]), x0=np.array([0.3, -1.0]),
                 costs=[dict(kind="trajectory", M=np.eye(2), p=np.array([0.1, 0.2]), w=np.array([0.5, 0.7])),
                        dict(kind="target", M=np.eye(2), p=np.ones(2), w=np.array([2.0, 3.0])),
                        dict(kind="control", N=np.eye(1), p=np.array([0.4]), w=np.array([0.1])),
                        dict(kind="mixed", M=np.ones((1, 2)), N=np.ones((1, 1)), p=np.array([0.3]), w=np.array([0.3]))],
                 constraints=[dict(kind="trajectory", E=np.eye(2), f=np.array([50.0, 40.0])),
                              dict(kind="mixed", E=np.array([[0.0, 1.0]]), G=np.array([[1.0]]), f=np.array([-0.5]), is_ineq=False),
                              dict(kind="control", G=np.eye(1), f=np.array([20.0])),
                              dict(kind="mixed", E=np.ones((1, 2)), G=np.ones((1, 1)), f=np.array([60.0])),
                              dict(kind="trajectory_bound", lower=np.array([-np.inf, -np.inf]), upper=np.array([2.0, 1.0])),
                              dict(kind="control_bound", lower=np.array([-1.0]), upper=np.array([1.5]))])
    yield "all_kinds_lmpc", mixed
    ist = dict(mixed)
    ist.update(initial_state=True, R=np.array([[2.0, 0.1], [0.1, 1.0]]), r=np.array([0.1, -0.2]),
               x0lb=np.array([0.3, -1.5]), x0ub=np.array([0.3, -0.5]))
    yield "all_kinds_initial_state", ist


def main():
    out = {}
    problems = {}
    for name, prob in small_cases():
        o = po.lmpc(prob)
        assert o["fail"] == 0, (name, o["fail"])
        for k in KEYS:
            out["%s/%s" % (name, k)] = np.asarray(o[k])
        out["%s/iter" % name] = np.asarray(o["iter"])
        problems[name] = json.loads(json.dumps(prob, default=lambda a: np.asarray(a).tolist()))
    np.savez_compressed(os.path.join(HERE, "small_cases.npz"), **out)
    json.dump(problems, open(os.path.join(HERE, "small_cases_problems.json"), "w"))
    # reference points of the C1 fixture: iteration counts, active rows, objective
    pts = {}
    for cost in ("target", "trajectory"):
        o = po.lmpc(wl.instance(wl.c1(cost), 0))
        pts[cost] = dict(iter=list(o["iter"]), nact=int(o["nact"]), objective=float(o["crval"]),
                         terminal_velocity=float(o["trajectory"][-1]))
    json.dump(pts, open(os.path.join(HERE, "c1_reference_points.json"), "w"), indent=1)
    print("wrote", sorted(problems), pts)


if __name__ == "__main__":
    main()
