"""Target for `ncu --set full --import-source on` of three GEMM launches of the ViT-B/16 B=256 step:
  1 out-proj forward  50432 x 768 x 768   +bias +residual   (short K: per-unit overheads show)
  2 FC1 forward       50432 x 3072 x 768  +bias, GELU and GELU' outputs
  3 FC2 forward       50432 x 768 x 3072  +bias +residual   (the dominant shape)
Two warm-up rounds, then one profiled: ncu -k regex:gemm_kernel -s 6 -c 3."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
from vit_pytorch_robust import _abi
dev = torch.device("cuda:0"); _abi.init(dev)
T, D, M3 = 50432, 768, 3072
def rnd(*s): return (torch.randn(*s, device=dev) / 8).to(torch.bfloat16)
x768, x3072 = rnd(T, D), rnd(T, M3)
w_fc1, w_fc2, w_out = rnd(M3, D), rnd(D, M3), rnd(D, D)
b768, b3072, res = torch.zeros(D, device=dev), torch.zeros(M3, device=dev), rnd(T, D)
o768 = torch.empty(T, D, device=dev, dtype=torch.bfloat16)
h, g = torch.empty(T, M3, device=dev, dtype=torch.bfloat16), torch.empty(T, M3, device=dev, dtype=torch.bfloat16)
for it in range(3):
    _abi.gemm(x768, w_out, o768, bias=b768, residual=res)
    _abi.gemm(x768, w_fc1, h, bias=b3072, epi=_abi.EPI_GELU_GRAD, out2=g)
    _abi.gemm(x3072, w_fc2, o768, bias=b768, residual=res)
torch.cuda.synchronize()
print("done")
