"""one lmpc_run of the default bench workload (C3 x 16384) with device-resident timing printed -- the command the ncu captures wrap"""
import sys
sys.path.insert(0, ".")
from copra_b200 import capi, workloads as wl
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
bp = wl.c1() if cfg == "c1" else wl.CONFIGS[cfg](batch=batch)
eng = capi.Engine(0)
for _ in range(reps):
    o = eng.lmpc_run(bp, want=("status", "iters"))
print(eng.last_solver(), eng.timing(), o["iters"].mean(axis=0), int((o["status"] != 0).sum()))
