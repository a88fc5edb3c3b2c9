import numpy as np, sys
sys.path.insert(0, ".")
from copra_b200 import capi
from oracle import pyoracle as po
T, mass = 0.005, 5.0
for N in (50, 100, 200, 300):
    A = np.array([[1.0, T], [0.0, 1.0]]); B = np.array([[0.5 * T * T / mass], [T / mass]]); c = np.array([(-9.81 / 2.0) * T ** 2, -9.81 * T])
    x0 = np.array([0.0, -5.0])
    prob = dict(name="pin", nx=2, nu=1, N=N, batch=1, A=A, B=B, d=c, x0=x0, initial_state=True,
        costs=[dict(kind="target", M=np.eye(2), p=np.zeros(2), w=np.array([10.0,100.0])), dict(kind="control", N=np.ones((1,1)), p=np.zeros(1), w=np.array([1e-2]))],
        constraints=[dict(kind="control_bound", lower=np.array([-np.inf]), upper=np.array([200.0]))],
        R=np.eye(2), r=np.array([3.0,-7.0]), x0lb=None, x0ub=None)
    eng = capi.Engine(0)
    out = eng.lmpc_run(prob)
    o = po.lmpc(prob)
    print(N, eng.last_solver(), "gpu status", out["status"], "iters", out["iters"], "nact", out["nact"], "| oracle", o["fail"], o["iter"], o["nact"], "x0 gpu", out["x"][0][:2], "xerr", np.abs(out["x"][0]-o["x"]).max())
    eng.close()
