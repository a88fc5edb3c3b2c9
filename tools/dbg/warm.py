import sys
import numpy as np
sys.path.insert(0, ".")
from copra_b200 import capi, workloads as wl
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.97
bp = wl.c3(batch=batch)
eng = capi.Engine(0)
out0 = eng.lmpc_run(bp, want=("status", "iters", "control", "iact"))
sizes = out0["sizes"]
print("full step", eng.timing()["total_ms"], "solve", eng.timing()["solve_ms"], "iters", out0["iters"].mean(axis=0))
x0 = np.asarray(bp["x0"]) * scale
cold = eng.lmpc_resolve(x0, sizes)
print("cold resolve total", eng.timing()["total_ms"], "solve", eng.timing()["solve_ms"], "iters", cold["iters"].mean(axis=0))
eng.set_warm_start(True)
eng.lmpc_run(bp, want=("status",))
warm = eng.lmpc_resolve(x0, sizes)
print("warm resolve total", eng.timing()["total_ms"], "solve", eng.timing()["solve_ms"], "iters", warm["iters"].mean(axis=0), "status!=0", int((warm["status"] != 0).sum()))
print("x diff warm vs cold: %.3e" % np.abs(warm["control"] - cold["control"]).max(), "same active sets:",
      all(set(a[a > 0]) == set(c[c > 0]) for a, c in zip(warm["iact"], cold["iact"])))
