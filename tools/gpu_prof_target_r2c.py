"""Target for `ncu --set full` of the kernels added in round 2 (session 3).  Per iteration, in this order:
  1 patch-embedding forward with the im2col in the TMA loads (ViT-B/16, B = 256)     gemm_kernel<256,1,0,0,1>
  2 its weight gradient (K = patch rows, image as the MN-major operand)                 gemm_kernel<256,1,0,0,1>
  3 general tcgen05 attention forward  (ViT-H/14: B = 32, 257 tokens, 16 heads, dh 80)  attn_fwd_big_kernel
  4 general tcgen05 attention backward (same shape)                                      attn_bwd_big_kernel
  5 the dominant GEMM (FC2 forward shape 50432 x 768 x 3072, plain store)               gemm_kernel<256,1,0,0,0>
  6 fused attention backward at the ViT-B/16 shape (B = 256, 197 tokens)                attn_bwd2_kernel
One warm-up iteration, one profiled: ncu --launch-skip 7 --launch-count 6 (iteration 1 has 7 launches: the fwd of item 6)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
from vit_pytorch_robust import _abi
dev = torch.device("cuda:0"); lib = _abi.init(dev)
sp = _abi.stream_ptr()
B, C, HW, P, D = 256, 3, 224, 16, 768
N = (HW // P) ** 2 + 1
img = torch.randn(B, C, HW, HW, device=dev).to(torch.bfloat16)
wp = (torch.randn(D, C * P * P, device=dev) / 28).to(torch.bfloat16)
bias = torch.zeros(D, device=dev); pos = torch.randn(N, D, device=dev)
xs = torch.empty(B * N, D, device=dev, dtype=torch.bfloat16)
dx = torch.randn(B * N, D, device=dev).to(torch.bfloat16); dw = torch.zeros(D, C * P * P, device=dev)
Bh, Nh, Hh, dh = 32, 257, 16, 80
qkv = torch.randn(Bh, Nh, 3 * Hh * dh, device=dev).to(torch.bfloat16); dout = torch.randn(Bh, Nh, Hh * dh, device=dev).to(torch.bfloat16)
out = torch.empty_like(dout); dqkv = torch.empty_like(qkv); lse = torch.empty(Bh, Hh, Nh, device=dev)
nb = lib.nrv_attn_bwd_workspace(Bh, Nh, Hh, dh); ws = torch.empty(nb, dtype=torch.uint8, device=dev)
T, M = 50432, 3072
a = torch.randn(T, M, device=dev).to(torch.bfloat16); w = (torch.randn(D, M, device=dev) / 32).to(torch.bfloat16)
o = torch.empty(T, D, device=dev, dtype=torch.bfloat16)
B2, N2, H2 = 256, 197, 12
qkv2 = torch.randn(B2, N2, 3 * H2 * 64, device=dev).to(torch.bfloat16); dout2 = torch.randn(B2, N2, H2 * 64, device=dev).to(torch.bfloat16)
out2 = torch.empty_like(dout2); dqkv2 = torch.empty_like(qkv2); lse2 = torch.empty(B2, H2, N2, device=dev)
nb2 = lib.nrv_attn_bwd_workspace(B2, N2, H2, 64); ws2 = torch.empty(max(nb2, 16), dtype=torch.uint8, device=dev)
_abi.check(lib.nrv_attn_fwd(qkv2.data_ptr(), out2.data_ptr(), lse2.data_ptr(), B2, N2, H2, 64, 0.125, 0, 0, 2, None, 0, sp))
for _ in range(2):
    _abi.check(lib.nrv_patch_embed_fwd(img.data_ptr(), B, C, HW, HW, P, P, wp.data_ptr(), C * P * P, bias.data_ptr(), pos.data_ptr(), D,
                                       N, 1, xs.data_ptr(), D, D, sp))
    _abi.check(lib.nrv_patch_embed_bwd_weight(img.data_ptr(), B, C, HW, HW, P, P, dx.data_ptr(), D, N, 1, dw.data_ptr(), C * P * P, D, sp))
    _abi.check(lib.nrv_attn_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), Bh, Nh, Hh, dh, dh ** -0.5, 0, 0, 2, None, 0, sp))
    _abi.check(lib.nrv_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), Bh, Nh, Hh, dh,
                                dh ** -0.5, 0, 0, 2, ws.data_ptr(), nb, sp))
    _abi.gemm(a, w, o)
    _abi.check(lib.nrv_attn_bwd(qkv2.data_ptr(), out2.data_ptr(), dout2.data_ptr(), lse2.data_ptr(), dqkv2.data_ptr(), B2, N2, H2, 64,
                                0.125, 0, 0, 2, ws2.data_ptr(), nb2, sp))
torch.cuda.synchronize()
print("ok")
