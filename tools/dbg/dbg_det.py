import numpy as np, sys, ctypes as C
sys.path.insert(0, ".")
from copra_b200 import capi, workloads as wl
which = sys.argv[1] if len(sys.argv) > 1 else "c5"
bp = wl.c5(batch=592) if which == "c5" else wl.c3(batch=2368)
eng = capi.Engine(0)
hb = capi.HostBatch(bp)
s = eng.sizes(hb)
B = bp["batch"]
def solve():
    out = dict(x=np.zeros((B, s["nvar"])), iters=np.zeros((B, 2), np.int32), status=np.zeros(B, np.int32))
    r = capi.Results(); r.memory = capi.HOST
    for k, v in out.items(): setattr(r, k, v.ctypes.data)
    eng._check(eng.lib.copra_b200_lmpc_solve(eng.h, C.byref(r)))
    return out
for build in range(2):
    eng.lmpc_build(hb)
    Q = eng.download(hb, "Q"); c = eng.download(hb, "c"); b = eng.download(hb, "bineq")
    if build == 0: Q0, c0, b0 = Q, c, b
    else: print("build stages identical:", np.array_equal(Q, Q0), np.array_equal(c, c0), np.array_equal(b, b0))
    runs = [solve() for _ in range(3)]
    for r in (1, 2):
        same = np.array_equal(runs[0]["x"], runs[r]["x"])
        diff = np.nonzero((runs[0]["x"] != runs[r]["x"]).any(1))[0]
        print("build", build, "solve 0 vs", r, "identical:", same, "differing instances:", diff[:10], [(runs[0]["iters"][i].tolist(), runs[r]["iters"][i].tolist(), float(np.abs(runs[0]["x"][i]-runs[r]["x"][i]).max())) for i in diff[:4]])
    if build == 0: first = runs[0]
    else: print("across builds identical:", np.array_equal(first["x"], runs[0]["x"]))
