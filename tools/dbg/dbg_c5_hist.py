import numpy as np, sys
sys.path.insert(0, ".")
from copra_b200 import capi, workloads as wl
bp = wl.c5(batch=1024)
eng = capi.Engine(0)
o = eng.lmpc_run(bp, want=("status", "iters", "nact"))
it = o["iters"].astype(np.int64)
passes = it[:, 0] + it[:, 1]
order = np.argsort(-passes)
print("top passes:", [(int(i), it[i].tolist(), int(o["nact"][i])) for i in order[:8]])
print("sum passes", passes.sum(), "mean", passes.mean(), "pct", np.percentile(passes, [50, 90, 99, 100]).tolist())
print(eng.timing())
