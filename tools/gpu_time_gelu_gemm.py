"""FC1 GEMM of ViT-B/16 B=256 (50432 x 3072 x 768) with the three epilogues: plain store, GELU (inference), GELU + GELU'
(training).  Times (CUDA events, 30 launches) and the error of h / g against the exact erf forms on fp64."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
from vit_pytorch_robust import _abi
dev = torch.device("cuda:0"); _abi.init(dev)
T, D, M = 50432, 768, 3072
torch.manual_seed(0)
x = torch.randn(T, D, device=dev).to(torch.bfloat16); w = (torch.randn(M, D, device=dev) / 16).to(torch.bfloat16)
b = torch.randn(M, device=dev) * 0.5 - 0.5
h = torch.empty(T, M, device=dev, dtype=torch.bfloat16); g = torch.empty_like(h)
def t(fn, n=30):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("plain   %.1f us" % t(lambda: _abi.gemm(x, w, h, bias=b)))
print("gelu    %.1f us" % t(lambda: _abi.gemm(x, w, h, bias=b, epi=_abi.EPI_GELU)))
print("gelu+g  %.1f us" % t(lambda: _abi.gemm(x, w, h, bias=b, epi=_abi.EPI_GELU_GRAD, out2=g)))
rows = slice(0, 4096)
u = (x[rows].double() @ w.double().t() + b.double())
u.requires_grad_(True)
ref_h = torch.nn.functional.gelu(u); ref_h.sum().backward(); ref_g = u.grad
for name, got, ref in (("h", h[rows], ref_h.detach()), ("g", g[rows], ref_g)):
    e = (got.double() - ref).abs()
    q = (ref.to(torch.bfloat16).double() - ref).abs()       # what rounding the exact value to bf16 costs
    print("%s: max abs err %.2e (bf16 rounding alone %.2e)  rms err %.3e (rounding alone %.3e)  rel-L2 %.3e (%.3e)" %
          (name, e.max(), q.max(), e.pow(2).mean().sqrt(), q.pow(2).mean().sqrt(), e.norm() / ref.norm(), q.norm() / ref.norm()))
neg = (u.detach() < -1.5)
e = (h[rows].double() - ref_h.detach()).abs()[neg]; q = (ref_h.detach().to(torch.bfloat16).double() - ref_h.detach()).abs()[neg]
print("h on u < -1.5: rms err %.3e (bf16 rounding alone %.3e), share of elements %.3f" % (e.pow(2).mean().sqrt(), q.pow(2).mean().sqrt(), neg.double().mean()))
