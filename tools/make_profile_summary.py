#!/usr/bin/env python
"""Regenerate profiles/r01_summary.md from the artefacts next to it:
  r01_bench_c2.json (bench.py line), r01_bench_c2_reference_arm.json, r01_bench_c{1,3,4,5}.json,
  r01_launches_c2.csv (ncu --metrics gpu__time_duration.sum launch list), r01_gi_small_c2_ncu_raw.txt (ncu --set full, raw page).
usage: python tools/make_profile_summary.py [round]"""
import collections
import csv
import json
import os
import sys

rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles")
P = lambda name: os.path.join(here, "%s_%s" % (rnd, name))

bench = json.load(open(P("bench_c2.json")))
ref = json.load(open(P("bench_c2_reference_arm.json"))) if os.path.exists(P("bench_c2_reference_arm.json")) else None
out = []
w = out.append
w("# Round 1 profile summary (B200, C2 = batch 4096 double-integrator LMPCs, N=50)\n")
w("Command: `python bench.py --steps %d --warmup %d` (gpurun, 1 GPU).  Raw line: `profiles/%s_bench_c2.json`.\n" % (bench["steps"], bench["warmup"], rnd))
w("| quantity | value |\n|---|---|")
w("| value (device-resident inputs) | %.0f solves/s, %.3f ms per %d-instance step |" % (bench["value"], bench["ms_per_step"], bench["instances"]))
w("| e2e (host buffers through the C ABI) | %.0f solves/s (H2D %d B + D2H %d B per step inside the timed region) |" % (
    bench["e2e"]["value"], bench["e2e"]["h2d_bytes_per_step"], bench["e2e"]["d2h_bytes_per_step"]))
if "cpu_baseline" in bench:
    w("| CPU baseline (oracle port, %d host threads) | %.0f solves/s |" % (bench["cpu_baseline"]["cores"], bench["cpu_baseline"]["value"]))
if ref:
    w("| reference arm (`--impl reference`, same port on all host threads) | %.0f solves/s |" % ref["value"])
w("| stage ms (library CUDA events) | %s |" % json.dumps({k: round(v, 4) for k, v in bench["stage_ms"].items()}))
r = bench["roofline"]
w("| K6 roofline (HBM, algorithmic bytes) | %.1f GB/s of %.1f measured = %.4f; ncu DRAM traffic per launch %s B |" % (r["achieved"], r["peak"], r["frac"], r["traffic"]))
f = r["fp64"]
w("| K6 FP64 | %.2f TFLOP/s algorithmic of %.1f measured DFMA peak (DMMA %.1f) = %.3f |" % (
    f["achieved_tflops"], f.get("measured_dfma_peak_tflops", 0.0), f.get("measured_dmma_peak_tflops", 0.0), f["frac"]))
w("| clocks | %s |" % json.dumps(bench["clocks"]))
w("")

# launch list
rows = [x for x in csv.reader(open(P("launches_c2.csv"))) if len(x) > 5]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for x in rows[1:]:
    try:
        v = float(x[iv].replace(",", ""))
    except ValueError:
        continue
    name = x[ik]
    if "chain_kernel" in name:  # the FP64 peak microbenchmark runs once, outside the timed region
        continue
    if "cb::" not in name and not name.startswith(("k1_", "k2_", "k3_", "k4_", "k7_", "gi_", "void k", "void gi", "void cb")):
        continue
    name = name.replace("cb::", "").replace("void ", "").split("(")[0]
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) / len(v) for v in agg.values())
w("## ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`, cold cache, serialised) -- shares of a step\n")
w("| kernel | launches | avg us | share |\n|---|---|---|---|")
k6_share = 0.0
for k, v in agg.items():
    m = sum(v) / len(v)
    w("| %s | %d | %.1f | %.1f%% |" % (k, len(v), m / 1000.0, 100.0 * m / tot))
    if k.startswith("gi_"):
        k6_share += m / tot
w("")
w("The bench's own stage split gives K6 a share of %.1f%% of the step; the ncu launch list gives %.1f%%: they agree.\n" % (
    100.0 * r["share_of_step"], 100.0 * k6_share))

# ncu raw
raw = {}
if os.path.exists(P("gi_small_c2_ncu_raw.txt")):
    for line in open(P("gi_small_c2_ncu_raw.txt")):
        if " = " in line and not line.startswith("#"):
            k, v = line.rsplit(" = ", 1)
            try:
                raw[k.split(" [")[0]] = float(v.replace(",", ""))
            except ValueError:
                pass
if raw:
    inst = raw.get("smsp__inst_executed.sum", 0.0)
    w("## gi_small_kernel, `ncu --set full` (%s_gi_small_c2_ncu_raw.txt, source hot spots in %s_gi_small_c2_source_hotspots.txt)\n" % (rnd, rnd))
    w("* %.3g warp instructions for %d instances = %.0f k per instance; issue slots %.1f%% busy, FP64 pipe %.1f%%, warps active %.1f%% of the SM's 64 "
      "(%d CTAs x 4 warps per SM; limits: registers %d, shared memory %d CTAs), %d registers/thread."
      % (inst, bench["config"]["batch_per_gpu"], inst / bench["config"]["batch_per_gpu"] / 1e3, raw.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0),
         raw.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 0), raw.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0),
         int(min(raw.get("launch__occupancy_limit_registers", 0), raw.get("launch__occupancy_limit_shared_mem", 0))),
         int(raw.get("launch__occupancy_limit_registers", 0)), int(raw.get("launch__occupancy_limit_shared_mem", 0)), int(raw.get("launch__registers_per_thread", 0))))
    w("* DRAM traffic %.1f MB read + %.1f MB written per launch (kernel time under ncu %.3f ms) vs %.1f MB algorithmic (Q + Aineq + vectors in, x + iact out): "
      "no re-reads; the kernel is instruction/latency bound, not bandwidth bound."
      % (raw.get("dram__bytes_read.sum", 0), raw.get("dram__bytes_write.sum", 0), raw.get("gpu__time_duration.sum", 0),
         r["achieved"] * 1e9 * bench["stage_ms"]["solve_ms"] * 1e-3 / 1e6))
notes = os.path.join(here, "%s_notes.md" % rnd)
if os.path.exists(notes):
    w("")
    w(open(notes).read().rstrip())
open(os.path.join(here, "%s_summary.md" % rnd), "w").write("\n".join(out) + "\n")
print("\n".join(out))
