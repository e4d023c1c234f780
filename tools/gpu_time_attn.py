"""CUDA-event timing of nrv_attn_fwd / nrv_attn_bwd (tcgen05 path) at the ViT-B/16 shape."""
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
from vit_pytorch_robust import _abi

dev = torch.device("cuda:0")
lib = _abi.init(dev)
for (B, N, H) in [(256, 197, 12), (1024, 64, 8), (128, 197, 16)]:
    dh = 64
    qkv = torch.randn(B, N, 3 * H * dh, device=dev).to(torch.bfloat16)
    dout = torch.randn(B, N, H * dh, device=dev).to(torch.bfloat16)
    out = torch.empty_like(dout)
    dqkv = torch.empty_like(qkv)
    lse = torch.empty(B, H, N, device=dev)
    nb = lib.nrv_attn_bwd_workspace(B, N, H, dh)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    sp = _abi.stream_ptr()
    def fwd():
        _abi.check(lib.nrv_attn_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, dh, dh ** -0.5, 0, 0, 2, None, 0, sp))
    def bwd():
        _abi.check(lib.nrv_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
                                    B, N, H, dh, dh ** -0.5, 0, 0, 2, ws.data_ptr(), nb, sp))
    for name, fn, mult in (("fwd", fwd, 4), ("bwd", bwd, 10)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = mult * B * H * N * N * dh
        print("attn %s B=%d N=%d H=%d: %.3f ms  %.0f TFLOP/s (algorithmic)" % (name, B, N, H, ms, fl / ms / 1e9), flush=True)
    # torch SDPA for context
    q, k, v = qkv.view(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
    for _ in range(3):
        torch.nn.functional.scaled_dot_product_attention(q, k, v)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        torch.nn.functional.scaled_dot_product_attention(q, k, v)
    e1.record()
    torch.cuda.synchronize()
    print("   torch SDPA fwd: %.3f ms" % (e0.elapsed_time(e1) / 10), flush=True)
