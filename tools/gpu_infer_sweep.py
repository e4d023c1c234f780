"""BASELINE.json configs[4]: ViT-H/14 224x224 bf16 inference-only sweep (eval() + no_grad), batch 1 .. 1024:
images/s and latency per batch, CUDA events, 1 B200."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
import vit_pytorch_robust as V  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
m = V.vit_h_14()
with torch.no_grad():
    m.heads.head.weight.normal_(std=0.02)
m = m.to(dev).eval()
F = 334.590e9  # forward FLOPs per image
for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024):
    x = torch.randn(B, 3, 224, 224, device=dev).to(torch.bfloat16)
    with torch.no_grad():
        for _ in range(3):
            y = m(x)
        torch.cuda.synchronize()
        n = 20 if B <= 64 else 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            y = m(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("ViT-H/14 inference B=%4d  %8.2f ms  %8.0f img/s  %6.0f model-TFLOP/s" % (B, ms, B / ms * 1e3, B * F / ms / 1e9), flush=True)
