import sys
sys.path.insert(0, ".")
from copra_b200 import capi, workloads as wl
bp = wl.c3(batch=int(sys.argv[1]) if len(sys.argv) > 1 else 2368)
eng = capi.Engine(0)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    o = eng.lmpc_run(bp, want=("status", "iters"))
print(eng.timing(), o["iters"].mean(axis=0))
