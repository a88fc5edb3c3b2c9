import numpy as np, sys, os
sys.path.insert(0, ".")
from copra_b200 import capi, workloads as wl
from oracle import pyoracle as po
bp = wl.c5(batch=592)
eng = capi.Engine(0)
outs = []
for rep in range(3):
    o = eng.lmpc_run(bp, want=("x", "status", "iters", "nact", "iact"))
    outs.append(o)
    it = o["iters"]
    print("rep", rep, eng.last_solver(), "status!=0:", int((o["status"] != 0).sum()), "max iters", it.max(0), "argmax", it[:, 0].argmax(), "mean", it.mean(0))
print("deterministic:", all(np.array_equal(outs[0][k], outs[r][k]) for r in (1, 2) for k in ("x", "iters", "iact")))
worst = np.argsort(-outs[0]["iters"][:, 0])[:6]
print("heaviest", worst, outs[0]["iters"][worst].tolist())
for i in list(worst[:2]) + [110]:
    o = po.lmpc(wl.instance(bp, int(i)))
    g = outs[0]
    print(i, "oracle", o["fail"], o["iter"], o["nact"], "| gpu", g["status"][i], g["iters"][i], g["nact"][i], "xerr", np.abs(g["x"][i] - o["x"]).max(),
          "same set", set(int(k) for k in o["iact"]) == set(int(k) for k in g["iact"][i] if k > 0))
