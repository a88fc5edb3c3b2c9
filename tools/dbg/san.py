"""small runs of every thin-solver kernel for compute-sanitizer: C3 (shared-factor form, cold + warm re-solve), C5 (cluster + hybrid queue)"""
import sys
import numpy as np
sys.path.insert(0, ".")
from copra_b200 import capi, workloads as wl
eng = capi.Engine(0)
which = sys.argv[1] if len(sys.argv) > 1 else "c3"
if which == "c3":
    bp = wl.c3(batch=40)
    o = eng.lmpc_run(bp)
    eng.set_warm_start(True)
    eng.lmpc_run(bp)
    w = eng.lmpc_resolve(np.asarray(bp["x0"]) * 0.9, o["sizes"])
    print("c3 ok", int((o["status"] != 0).sum()), int((w["status"] != 0).sum()), w["iters"].mean(axis=0))
else:
    bp = wl.c5(batch=20)
    o = eng.lmpc_run(bp)
    print("c5 ok", eng.last_solver(), int((o["status"] != 0).sum()), o["iters"].mean(axis=0))
