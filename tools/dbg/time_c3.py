import sys, time
sys.path.insert(0, ".")
from copra_b200 import capi, workloads as wl
bp = wl.c3(batch=int(sys.argv[1]) if len(sys.argv) > 1 else 4736)
eng = capi.Engine(0)
best = 1e9
for _ in range(4):
    o = eng.lmpc_run(bp, want=("status",))
    best = min(best, eng.timing()["solve_ms"])
print("solve_ms %.2f  -> %.0f solves/s (solve only)" % (best, bp["batch"] / best * 1e3 if isinstance(bp, dict) else 0))
