"""Data-parallel sharder (copra_b200_multi_*): the sharded result must be BITWISE the single-device result,
because instances are independent and every shard runs the same kernels.  With fewer than two GPUs the two shards share
device 0 (two handles, two streams, two host threads), which exercises the same host logic."""
import numpy as np
import pytest

from copra_b200 import capi, workloads as wl

pytestmark = pytest.mark.gpu

KEYS = ("control", "trajectory", "x", "status", "iters", "nact", "iact")


def _devices(k=2):
    n = capi.load().copra_b200_device_count()
    return list(range(k)) if n >= k else [0] * k


@pytest.mark.parametrize("make", [lambda: wl.c2(batch=301), lambda: wl.c3(batch=37), lambda: wl.c4(batch=130)])
def test_sharded_run_is_bitwise_single_device(engine, make):
    bp = make()
    hb = capi.HostBatch(bp)
    sizes = engine.sizes(hb)
    ref = engine.lmpc_run(hb)
    me = capi.MultiEngine(_devices(2))
    try:
        out = me.lmpc_run(hb, sizes)
        shards = me.shards()
    finally:
        me.close()
    per = -(-bp["batch"] // 2)
    assert [(lo, hi) for _, lo, hi in shards] == [(0, per), (per, bp["batch"])]
    assert (ref["status"] == 0).all()
    for k in KEYS:
        assert np.array_equal(out[k], ref[k]), (bp["name"], k)


def test_three_shards_and_empty_shard(engine):
    bp = wl.c2(batch=5)
    hb = capi.HostBatch(bp)
    sizes = engine.sizes(hb)
    ref = engine.lmpc_run(hb)
    me = capi.MultiEngine(_devices(3) + [0])  # 4 shards of ceil(5/4) = 2: [0,2) [2,4) [4,5) and an empty one
    try:
        out = me.lmpc_run(hb, sizes)
        assert [(lo, hi) for _, lo, hi in me.shards()] == [(0, 2), (2, 4), (4, 5), (5, 5)]
        assert me.launch_count() > 0
    finally:
        me.close()
    for k in KEYS:
        assert np.array_equal(out[k], ref[k]), k


def test_sharded_resolve_matches_single_device(engine):
    bp = wl.c2(batch=64)
    hb = capi.HostBatch(bp)
    sizes = engine.sizes(hb)
    x0 = np.array(bp["x0"]) * np.array([1.0, 0.9])
    engine.lmpc_run(hb)
    ref = engine.lmpc_resolve(x0, sizes)
    me = capi.MultiEngine(_devices(2))
    try:
        me.lmpc_run(hb, sizes)
        out = me.lmpc_resolve(x0, sizes)
    finally:
        me.close()
    for k in KEYS:
        assert np.array_equal(out[k], ref[k]), k


def test_device_arrays_are_rejected(engine):
    bp = wl.c2(batch=4)
    hb = capi.HostBatch(bp)
    sizes = engine.sizes(hb)
    me = capi.MultiEngine(_devices(2))
    try:
        hb.problem.memory = capi.DEVICE
        with pytest.raises(capi.CopraB200Error):
            me.lmpc_run(hb, sizes)
    finally:
        hb.problem.memory = capi.HOST
        me.close()
