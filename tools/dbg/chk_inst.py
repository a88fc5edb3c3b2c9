import sys
import numpy as np
sys.path.insert(0, ".")
from copra_b200 import capi, workloads as wl
from oracle import pyoracle as po
cfg, batch, inst = sys.argv[1], int(sys.argv[2]), [int(a) for a in sys.argv[3:]]
bp = wl.CONFIGS[cfg](batch=batch)
eng = capi.Engine(0)
o = eng.lmpc_run(bp, want=("status", "iters", "control", "iact"))
for i in inst:
    r = po.lmpc(wl.instance(bp, i))
    e = np.abs(o["control"][i] - r["control"]).max() / max(1.0, np.abs(r["control"]).max())
    print("inst %d iters %s x err %.3e |x|max %.3g" % (i, o["iters"][i].tolist(), e, np.abs(r["control"]).max()))
