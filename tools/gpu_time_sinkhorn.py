"""Sinkhorn (robust=True) attention at the ViT-B/16 B=256 shape: tcgen05 forward vs CUDA-core forward, backward, and the
softmax kernels for scale."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
from vit_pytorch_robust import _abi
dev = torch.device("cuda:0"); lib = _abi.init(dev); sp = _abi.stream_ptr()
B, N, H, dh = 256, 197, 12, 64
qkv = torch.randn(B, N, 3 * H * dh, device=dev).to(torch.bfloat16); dout = torch.randn(B, N, H * dh, device=dev).to(torch.bfloat16)
out = torch.empty_like(dout); dqkv = torch.empty_like(qkv)
stats = torch.empty(B, H, 8, N, device=dev)
nb = lib.nrv_attn_bwd_workspace(B, N, H, dh); ws = torch.empty(nb, dtype=torch.uint8, device=dev)
def t(fn, n=5):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def fwd(mode, impl): _abi.check(lib.nrv_attn_fwd(qkv.data_ptr(), out.data_ptr(), stats.data_ptr(), B, N, H, dh, dh ** -0.5, mode, 0, impl, None, 0, sp))
def bwd(mode, impl): _abi.check(lib.nrv_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), stats.data_ptr(), dqkv.data_ptr(), B, N, H, dh, dh ** -0.5, mode, 0, impl, ws.data_ptr(), nb, sp))
print("softmax  fwd tcgen05 %.3f ms | bwd tcgen05 %.3f ms" % (t(lambda: fwd(0, 2)), t(lambda: bwd(0, 2))))
print("sinkhorn fwd tcgen05 %.3f ms | fwd CUDA cores %.3f ms" % (t(lambda: fwd(1, 2)), t(lambda: fwd(1, 1))))
fwd(1, 2)
print("sinkhorn bwd tcgen05 %.3f ms | bwd CUDA cores %.3f ms" % (t(lambda: bwd(1, 2)), t(lambda: bwd(1, 1))))
