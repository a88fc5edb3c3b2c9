#!/usr/bin/env python
"""Times the FP64 tensor-core GEMM (copra_b200_dgemm_batch) on device-resident operands; used for the ncu captures of
the dense assembly path.  usage: python tools/gemm_probe.py [M N K batch transA]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from copra_b200 import capi  # noqa: E402

M, N, K, batch, ta = (int(v) for v in (sys.argv[1:6] + ["602", "300", "602", "256", "0"][len(sys.argv) - 1:]))
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
eng = capi.Engine(0, stream=stream.cuda_stream)
ar, ac = (K, M) if ta else (M, K)
A = torch.randn(batch, ac, ar, dtype=torch.float64, device=dev)  # column-major per instance
B = torch.randn(batch, N, K, dtype=torch.float64, device=dev)
Cc = torch.zeros(batch, N, M, dtype=torch.float64, device=dev)


def run():
    rc = eng.lib.copra_b200_dgemm_batch(eng.h, ta, M, N, K, 1.0, A.data_ptr(), ar, ar * ac, B.data_ptr(), K, K * N, 0.0,
                                        Cc.data_ptr(), M, M * N, batch, capi.DEVICE)
    assert rc == 0


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(10):
    run()
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
ref = torch.matmul(A.transpose(1, 2) if not ta else A, B.transpose(1, 2)) if False else None
flops = 2.0 * M * N * K * batch
print("dgemm %dx%dx%d batch %d transA=%d: %.3f ms  %.2f TFLOP/s" % (M, N, K, batch, ta, ms, flops / ms / 1e9))
# correctness spot check against torch
Am = A.transpose(1, 2) if not ta else A  # logical (M,K) or stored (K,M)->op gives A^T
want = torch.matmul(Am if not ta else A, B.transpose(1, 2))
got = Cc.transpose(1, 2)
print("max abs err vs torch.matmul: %.3e" % (got - want).abs().max().item())
