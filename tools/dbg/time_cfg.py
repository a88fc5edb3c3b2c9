"""usage: time_cfg.py <config> <batch> [oracle-check count]: solve time of one config + parity of a few instances with the oracle"""
import sys
import numpy as np
sys.path.insert(0, ".")
from copra_b200 import capi, workloads as wl
cfg, batch = sys.argv[1], int(sys.argv[2])
ncheck = int(sys.argv[3]) if len(sys.argv) > 3 else 0
bp = wl.c1() if cfg == "c1" else wl.CONFIGS[cfg](batch=batch)
eng = capi.Engine(0)
best = 1e9
for _ in range(3):
    o = eng.lmpc_run(bp, want=("status", "iters", "control", "iact"))
    best = min(best, eng.timing()["solve_ms"])
print(cfg, "solver:", eng.last_solver() if hasattr(eng, "last_solver") else "?", "solve_ms %.2f -> %.0f solves/s (solve only)" % (best, batch / best * 1e3),
      "status!=0:", int((o["status"] != 0).sum()), "iters mean", o["iters"].mean(axis=0))
if ncheck:
    from oracle import pyoracle as po
    it = o["iters"].astype(np.int64).sum(axis=1)
    idx = list(np.argsort(-it)[:ncheck // 2]) + list(range(ncheck - ncheck // 2))
    worst = 0.0
    for i in idx:
        r = po.lmpc(wl.instance(bp, int(i)))
        e = np.abs(o["control"][i] - r["control"]).max() / max(1.0, np.abs(r["control"]).max())
        same = set(int(k) for k in o["iact"][i] if k > 0) == set(int(k) for k in r["iact"])
        worst = max(worst, e)
        print("  inst %d iters gpu %s oracle %s  x err %.2e  active set %s" % (i, o["iters"][i].tolist(), list(r["iter"]) if "iter" in r else "?", e, "same" if same else "DIFFERENT"))
    print("worst x err %.2e" % worst)
