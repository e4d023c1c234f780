"""Sustained (power-capped) GEMM throughput: libnrvit's tcgen05 GEMM vs cuBLAS on the ViT-B/16 step's shapes.

Short bursts run at boost clocks; a training step runs for seconds at the 1 kW cap.  This loops the step's
forward/dX GEMM shapes (plain-store epilogue on both sides) for ~2 s each and reports TFLOP/s under load, which is
the number to compare with MEASURED_PEAKS.json's bf16_tflops_sustained.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
from vit_pytorch_robust import _abi  # noqa: E402

dev = torch.device("cuda:0")
T = 256 * 197
SHAPES = [("qkv", T, 2304, 768), ("out", T, 768, 768), ("fc1", T, 3072, 768), ("fc2", T, 768, 3072)]
bufs = []
for name, M, N, K in SHAPES:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    b = torch.randn(N, K, device=dev).to(torch.bfloat16)
    o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    bufs.append((name, M, N, K, a, b, o))
flops_round = sum(2.0 * M * N * K for _, M, N, K in SHAPES)


def run_ours():
    for name, M, N, K, a, b, o in bufs:
        _abi.gemm(a, b, o)


def run_cublas():
    for name, M, N, K, a, b, o in bufs:
        torch.matmul(a, b.t(), out=o)


def sustained(fn, seconds=2.5):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    # calibrate
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / 20
    n = max(20, int(seconds * 1e3 / per))
    # heat-up half, then measure the second half
    for _ in range(n // 2):
        fn()
    e0.record()
    for _ in range(n // 2):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (n // 2)
    return flops_round / ms / 1e9, ms


for label, fn in (("cuBLAS", run_cublas), ("libnrvit", run_ours), ("cuBLAS", run_cublas), ("libnrvit", run_ours)):
    tf, ms = sustained(fn)
    print("%-9s sustained over the four forward shapes: %7.1f TFLOP/s  (%.3f ms per round)" % (label, tf, ms), flush=True)
    time.sleep(0.5)
