import numpy as np, sys, os
sys.path.insert(0, ".")
from copra_b200 import capi, workloads as wl
which = sys.argv[1] if len(sys.argv) > 1 else "c5"
bp = wl.c5(batch=3) if which == "c5" else wl.c3(batch=20)
eng = capi.Engine(0)
o = eng.lmpc_run(bp, want=("x", "status", "iters", "nact", "iact"))
print(which, eng.last_solver(), o["status"], o["iters"].tolist())
