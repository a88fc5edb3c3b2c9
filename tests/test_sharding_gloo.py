"""N>1 host logic on CPU: world_size-2 gloo run of the instance-index sharding + final gather that
bench.py uses (no collective on the hot path; only per-rank summaries cross ranks)."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from copra_b200 import workloads as wl
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    full = wl.c2(batch=10)
    mine, (lo, hi) = wl.shard(full, rank, world)
    assert mine["batch"] == hi - lo
    # every rank "solves" its shard: here a per-instance checksum of the parameters stands in for the result
    local = torch.tensor(np.asarray(mine["x0"])[:, 1] * 2.0 + np.asarray(mine["costs"][0]["p"])[:, 1])
    pad = torch.zeros(5, dtype=torch.float64); pad[: local.numel()] = local
    gathered = [torch.zeros(5, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, pad)                      # the final gather
    t = torch.tensor([float(rank + 1)]); dist.all_reduce(t, op=dist.ReduceOp.MAX)   # max-over-ranks timing
    if rank == 0:
        got = torch.cat([g[: (min(10, (r + 1) * 5) - min(10, r * 5))] for r, g in enumerate(gathered)]).numpy()
        want = np.asarray(full["x0"])[:, 1] * 2.0 + np.asarray(full["costs"][0]["p"])[:, 1]
        assert np.array_equal(got, want), (got, want)
        assert t.item() == world
        # weak-scaling batches use a different seed per rank (bench.py make_batch)
        assert not np.array_equal(wl.c2(batch=4, seed_offset=0)["x0"], wl.c2(batch=4, seed_offset=1)["x0"])
        print("GLOO_OK")
    dist.barrier(); dist.destroy_process_group()
""") % ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_shard_and_gather_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GLOO_OK" in r.stdout


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` (CPU oracle port) prints the contract's JSON line without a GPU"""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--batch", "64"], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "solves/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
