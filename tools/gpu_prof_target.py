"""Small target for `ncu --set full`: per iteration, in this order, the dominant GEMM (FC2 forward shape, plain
store), the FC1 shape with a plain store, the FC1 GELU-epilogue GEMM (two outputs), the FC2-dX GEMM with the
multiply epilogue, and the attention forward / backward kernels at the ViT-B/16 B=256 shapes."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
from vit_pytorch_robust import _abi
dev = torch.device("cuda:0"); lib = _abi.init(dev)
T, D, M = 50432, 768, 3072
a = torch.randn(T, M, device=dev).to(torch.bfloat16); w = (torch.randn(D, M, device=dev) / 32).to(torch.bfloat16)
o = torch.empty(T, D, device=dev, dtype=torch.bfloat16)
x = torch.randn(T, D, device=dev).to(torch.bfloat16); w1 = (torch.randn(M, D, device=dev) / 16).to(torch.bfloat16)
h = torch.empty(T, M, device=dev, dtype=torch.bfloat16); g = torch.empty_like(h); b1 = torch.zeros(M, device=dev)
B, N, H, dh = 256, 197, 12, 64
qkv = torch.randn(B, N, 3 * H * dh, device=dev).to(torch.bfloat16); dout = torch.randn(B, N, H * dh, device=dev).to(torch.bfloat16)
out = torch.empty_like(dout); dqkv = torch.empty_like(qkv); lse = torch.empty(B, H, N, device=dev)
nb = lib.nrv_attn_bwd_workspace(B, N, H, dh); ws = torch.empty(nb, dtype=torch.uint8, device=dev)
sp = _abi.stream_ptr()
for _ in range(3):
    _abi.gemm(a, w, o)
    _abi.gemm(x, w1, h)
    _abi.gemm(x, w1, h, bias=b1, epi=_abi.EPI_GELU_GRAD, out2=g)
    _abi.gemm(o, w, h, b_layout=_abi.NRV_MN_MAJOR, epi=_abi.EPI_MUL, aux=g)
    _abi.check(lib.nrv_attn_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, dh, dh ** -0.5, 0, 0, 2, None, 0, sp))
    _abi.check(lib.nrv_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), B, N, H, dh, dh ** -0.5, 0, 0, 2, ws.data_ptr(), nb, sp))
torch.cuda.synchronize()
print("ok")
