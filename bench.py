#!/usr/bin/env python
"""bench.py -- batched LMPC solves/sec (build + QP) on B200, the metric BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2] [--batch B] [--impl reference]

A *step* is one pass of the hot path (K1 condense -> K2-K4 assemble -> K6 Goldfarb-Idnani solve ->
K7 rollout) over one batch of synthetic controllers.  At N GPUs every rank owns one GPU and its own
`batch` instances (weak scaling, sharded by instance index, no collective on the hot path; the only
cross-GPU traffic is the final gather of per-rank timings/status counts).

`value`  : whole-job solves/s with the per-instance parameters already resident in HBM, timed with
           CUDA events per step (L2 flushed between steps), max over ranks.
`e2e`    : the same metric through the C ABI with HOST buffers (pinned): H2D of the parameters and
           D2H of control/trajectory/status inside the timed region.
`roofline`, `cpu_baseline`: see DESIGN.md "Measurement".
`--impl reference` times the CPU oracle port of copra's Eigen + eigen-quadprog path (the reference
cannot be built offline) on all host threads, same config/metric.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from copra_b200 import workloads as wl  # noqa: E402

METRIC = "batched LMPC solves/sec (build+QP)"
UNIT = "solves/s"
DEFAULT_BATCH = dict(c1=1, c2=4096, c3=16384, c4=8192, c5=1024)
WORKLOAD_NAME = dict(
    c1="C1: single double-integrator LMPC (tests/systems.h BoundedSystem, N=300), batch 1",
    c2="C2: batch 4096 double-integrator LMPCs, horizon 50, trajectory/control bounds, FP64",
    c3="C3: walking CoM preview (3rd-order LIPM, 6 states/2 jerk controls, horizon 160, ZMP mixed inequalities)",
    c4="C4: InitialStateLMPC on the double integrator, target + control costs",
    c5="C5: condensing stress nx=12 nu=4 horizon 200")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def k6_algorithmic_bytes(sz):
    n, meq, m, q = sz["nvar"], sz["meq"], sz["mineq"], sz["q"]
    return 8 * (n * n + n + (meq + m) * n + (meq + m) + 2 * n) + 8 * n + 4 * q + 12


def k6_algorithmic_flops(sz, iters):
    """F_alg = 2/3 n^3 + I (10 n^2 + 2 n q), I = outer iterations"""
    n, q = sz["nvar"], sz["q"]
    return (2.0 / 3.0) * n ** 3 + iters * (10.0 * n * n + 2.0 * n * q)


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the K6 kernel from the committed
# `ncu --set full` capture (profiles/r01_gi_small_c2_ncu_raw.txt); only valid for that exact workload
NCU_TRAFFIC = {("c2", 4096): 175565056 + 7223296}
FP64_PEAK_TFLOPS = 37.0  # nominal B200 FP64 (vector == tensor); not in MEASURED_PEAKS.json


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            rows = [r.strip().split(", ") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
            sm = [float(r[1]) for r in rows if len(r) >= 9]
            power = [float(r[3]) for r in rows if len(r) >= 9]
            if sm:
                # "under load": samples in the upper half of the power range
                thr = (min(power) + max(power)) / 2.0 if max(power) - min(power) > 50 else -1
                load = [s for s, p in zip(sm, power) if p >= thr] or sm
                out["sm_mhz"] = statistics.median(load)
                out["sm_max_mhz"] = float(rows[0][2])
                out["samples"] = len(sm)
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for k, nm in enumerate(names):
                    if any(r[5 + k].strip().lower().startswith("active") for r in rows if len(r) >= 9):
                        out["reasons"].append(nm)
        except Exception:
            pass
        return out


def make_batch(config, batch, rank):
    fn = wl.CONFIGS[config]
    if config == "c1":
        return fn()
    return fn(batch=batch, seed_offset=rank)


def cpu_sample(bp, max_instances, reps=3, threads=None):
    """oracle port on the host cores, one instance per thread"""
    from oracle import pyoracle as po
    n = min(bp["batch"], max_instances)
    probs = [wl.instance(bp, i) for i in range(n)]
    best = None
    for _ in range(reps):
        r = po.lmpc_batch(probs, threads=threads)
        if best is None or r["wall"] < best["wall"]:
            best = r
    return dict(value=n / best["wall"], cores=best["threads"], n=n, wall=best["wall"],
                p50_ms=float(np.median(best["t_inst"]) * 1e3), ok=int((best["fail"] == 0).sum()))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    config = args.config
    batch = args.batch or DEFAULT_BATCH[config]
    bp = make_batch(config, batch, 0)
    # bounded sample: ~10-30 s of CPU work per step at most
    per_inst = dict(c1=0.06, c2=0.0008, c3=0.18, c4=0.0012, c5=1.5)[config]
    from oracle import pyoracle as po
    cores = po.hw_threads()
    sample = int(max(cores, min(batch, (15.0 * cores) / per_inst / max(1, args.steps + args.warmup))))
    sample = min(sample, batch) if config != "c1" else 1
    probs = [wl.instance(bp, i) for i in range(sample)]
    for _ in range(args.warmup):
        po.lmpc_batch(probs[:max(1, min(sample, cores))])
    times = []
    for _ in range(args.steps):
        r = po.lmpc_batch(probs)
        times.append(r["wall"])
    t = sum(times)
    value = sample * args.steps / t
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * t / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic", impl="reference",
                config=dict(workload=WORKLOAD_NAME[config], config=config, batch_per_step=sample,
                            note="CPU restatement (oracle port) of copra's Eigen + eigen-quadprog path; the reference "
                                 "itself cannot be built offline (no Eigen/eigen-quadprog/gfortran)"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port",
                                  sample="%d instances of %s per step, one instance per host thread" % (sample, config)),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c2", choices=sorted(wl.CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU per step (default: the config's batch)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from copra_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    config = args.config
    batch = args.batch or DEFAULT_BATCH[config]
    bp = make_batch(config, batch, rank)
    batch = bp["batch"]

    stream = torch.cuda.Stream(device=dev)  # a real (non-legacy) stream shared by torch events and the engine
    torch.cuda.set_stream(stream)
    eng = capi.Engine(local, stream=stream.cuda_stream)
    hb_dev = capi.HostBatch(bp, device_tensors=dev)
    hb_host = capi.HostBatch(bp, pin=True)
    sz = eng.sizes(hb_dev)

    # device-resident results (value path) and pinned host results (e2e path)
    d_control = torch.empty(batch * sz["nU"], dtype=torch.float64, device=dev)
    d_traj = torch.empty(batch * sz["X"], dtype=torch.float64, device=dev)
    d_status = torch.empty(batch, dtype=torch.int32, device=dev)
    d_iters = torch.empty(2 * batch, dtype=torch.int32, device=dev)
    d_nact = torch.empty(batch, dtype=torch.int32, device=dev)
    d_iact = torch.empty(batch * sz["nvar"], dtype=torch.int32, device=dev)
    rd = capi.Results()
    rd.memory = capi.DEVICE
    rd.control, rd.trajectory, rd.status = d_control.data_ptr(), d_traj.data_ptr(), d_status.data_ptr()
    rd.iters, rd.nact, rd.iact = d_iters.data_ptr(), d_nact.data_ptr(), d_iact.data_ptr()

    h_control = torch.empty(batch * sz["nU"], dtype=torch.float64).pin_memory()
    h_traj = torch.empty(batch * sz["X"], dtype=torch.float64).pin_memory()
    h_status = torch.empty(batch, dtype=torch.int32).pin_memory()
    rh = capi.Results()
    rh.memory = capi.HOST
    rh.control, rh.trajectory, rh.status = h_control.data_ptr(), h_traj.data_ptr(), h_status.data_ptr()
    d2h_bytes = h_control.numel() * 8 + h_traj.numel() * 8 + h_status.numel() * 4

    import ctypes as C
    lib = eng.lib

    def step_device():
        rc = lib.copra_b200_lmpc_run(eng.h, C.byref(hb_dev.problem), C.byref(rd))
        if rc:
            eng._check(rc)

    def step_host():
        rc = lib.copra_b200_lmpc_run(eng.h, C.byref(hb_host.problem), C.byref(rh))
        if rc:
            eng._check(rc)

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up -------------------------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()  # samples clocks / throttle reasons through warm-up, the timed region and the e2e loop
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    n_ok = int((d_status == 0).sum().item())

    # ---- value: device-resident inputs, CUDA events per step, L2 flushed between steps ----------
    l0 = eng.launch_count()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stage = dict(condense_ms=0.0, assemble_ms=0.0, solve_ms=0.0, rollout_ms=0.0)
    barrier()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        starts[k].record(stream)
        step_device()
        stops[k].record(stream)
    barrier()
    wall1 = time.perf_counter()
    launches = eng.launch_count() - l0
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    # per-stage device times of one more (untimed) step, from the library's own CUDA events
    flush.fill_(1)
    step_device()
    tm = eng.timing()
    for k in stage:
        stage[k] = tm[k]

    total_ms = sum(step_ms)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * batch * args.steps / (total_ms_max * 1e-3)

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region -----------------
    for _ in range(2):
        step_host()
    e2e_steps = args.steps
    barrier()
    e0 = time.perf_counter()
    for k in range(e2e_steps):
        step_host()
    torch.cuda.synchronize()
    e1 = time.perf_counter()
    te = torch.tensor([e1 - e0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * batch * e2e_steps / float(te.item())
    ok_host = int((h_status == 0).sum().item())
    clocks = sampler.stop()

    # ---- final gather of per-rank results (status counts, timings): the only cross-GPU traffic ---
    summary = torch.tensor([float(n_ok), float(ok_host), total_ms, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        gathered = [torch.empty_like(summary) for _ in range(world)]
        dist.all_gather(gathered, summary)
        summary_all = torch.stack(gathered).cpu().numpy()
    else:
        summary_all = summary.cpu().numpy()[None, :]

    if rank == 0:
        peak, peak_src = peaks()
        dfma_tf, dmma_tf = eng.fp64_peaks()  # measured on this device, outside the timed region
        k6_bytes = k6_algorithmic_bytes(sz) * batch
        mean_iters = float(d_iters.view(-1, 2)[:, 0].double().mean().item())
        k6_flops = k6_algorithmic_flops(sz, mean_iters) * batch
        solve_s = stage["solve_ms"] * 1e-3
        achieved = k6_bytes / solve_s / 1e9 if solve_s > 0 else 0.0
        step_total = stage["condense_ms"] + stage["assemble_ms"] + stage["solve_ms"] + stage["rollout_ms"]
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=total_ms_max / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="f64", data="synthetic",
            config=dict(workload=WORKLOAD_NAME[config], config=config, batch_per_gpu=batch, nx=bp["nx"], nu=bp["nu"],
                        horizon=bp["N"], n_vars=sz["nvar"], n_ineq_rows=sz["mineq"], q_rows=sz["q"],
                        parallelism="instance-sharded x%d, no collective on the hot path" % world,
                        l2="flushed between timed steps (256 MiB fill)"),
            p50_ms_per_batch=float(np.median(step_ms)), p50_us_per_solve_amortised=float(np.median(step_ms)) * 1e3 / batch,
            solved_ok=int(summary_all[:, 0].sum()), instances=world * batch,
            stage_ms=stage,
            roofline=dict(kernel=("gi_small_kernel" if sz["nvar"] <= 64 else "gi_cluster_kernel / gi_batch_kernel") + " (K5+K6)", bound="hbm", achieved=achieved, peak=peak, unit="GB/s",
                          frac=achieved / peak, traffic=NCU_TRAFFIC.get((config, batch)), peak_source=peak_src,
                          fp64=dict(achieved_tflops=(k6_flops / solve_s / 1e12) if solve_s > 0 else 0.0, measured_dfma_peak_tflops=dfma_tf, measured_dmma_peak_tflops=dmma_tf,
                                    nominal_peak_tflops=FP64_PEAK_TFLOPS,
                                    frac=(k6_flops / solve_s / 1e12 / dfma_tf) if solve_s > 0 and dfma_tf > 0 else 0.0, mean_outer_iterations=mean_iters,
                                    note="algorithmic flops 2/3 n^3 + I(10 n^2 + 2 n q) vs the DFMA-chain peak measured by copra_b200_fp64_peaks on this device"),
                          share_of_step=(stage["solve_ms"] / step_total) if step_total > 0 else None,
                          note="K6 is FP64-latency bound by arithmetic intensity; HBM fraction reported as north_star asks"),
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=hb_host.h2d_bytes, d2h_bytes_per_step=d2h_bytes,
                     solved_ok=int(summary_all[:, 1].sum())),
            gpu_launches=int(summary_all[:, 3].sum()),
            clocks=dict(sm_mhz=clocks["sm_mhz"], sm_max_mhz=clocks["sm_max_mhz"], reasons=clocks["reasons"],
                        samples=clocks["samples"]),
            wall_s_timed_region=wall1 - wall0)
        if not args.no_cpu_baseline and world == 1:
            per_inst = dict(c1=0.06, c2=0.0008, c3=0.18, c4=0.0012, c5=1.5)[config]
            from oracle import pyoracle as po
            cores = po.hw_threads()
            nmax = int(max(cores, min(batch, 5.0 * cores / per_inst)))
            cb = cpu_sample(bp, nmax, reps=3)
            line["cpu_baseline"] = dict(value=cb["value"], unit=UNIT, cores=cb["cores"], kind="port",
                                        sample="%d instances of %s, best of 3, one instance per host thread; "
                                               "p50 %.3f ms per solve" % (cb["n"], config, cb["p50_ms"]))
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
