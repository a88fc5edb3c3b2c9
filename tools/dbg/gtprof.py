"""Aggregate the clock64 phase timers of the thin solver (build with `make EXTRA=-DGT_PROFILE`): cycles per outer iteration."""
import re, subprocess, sys, os
sys.path.insert(0, ".")
code = r'''
import sys
sys.path.insert(0, ".")
from copra_b200 import capi, workloads as wl
bp = wl.c3(batch=int(sys.argv[1]))
eng = capi.Engine(0)
o = eng.lmpc_run(bp, want=("status", "iters"))
print(eng.timing())
'''
batch = sys.argv[1] if len(sys.argv) > 1 else "2368"
out = subprocess.run([sys.executable, "-c", code, batch], capture_output=True, text=True).stdout
names = ["init", "prod", "sel", "normal+Jt'a", "q1", "Jz+S", "step+add", "drop"]
tot = dict.fromkeys(names, 0); iters = 0; cnt = 0
for ln in out.splitlines():
    if not ln.startswith("GTPROF"):
        if "solve_ms" in ln: print(ln)
        continue
    m = re.search(r"iters=(\d+) drops=(\d+)", ln)
    iters += int(m.group(1)); cnt += 1
    for nm in names:
        tot[nm] += int(re.search(re.escape(nm) + r" (\d+)", ln).group(1))
print("instances", cnt, "mean iters", iters / max(cnt, 1))
for nm in names:
    print(f"{nm:14s} {tot[nm] / max(iters, 1):10.0f} cycles / outer iteration")
print("sum", sum(tot.values()) / max(iters, 1))
