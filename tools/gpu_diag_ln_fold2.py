import os, sys, torch
sys.path.insert(0, "noise-robust-vit_b200")
from vit_pytorch_robust import _abi
dev = torch.device("cuda:0"); lib = _abi.init(dev); sp = _abi.stream_ptr()
def rel(a, b): return ((a.double() - b.double()).norm() / b.double().norm()).item()
torch.manual_seed(0)
M, N, K = 8192, 2304, 768
x = (torch.randn(M, K, device=dev) * 1.5 + 0.05).to(torch.bfloat16)
a = (6.0 / (K + N)) ** 0.5
W32 = (torch.rand(N, K, device=dev) * 2 - 1) * a
W = W32.to(torch.bfloat16)
gamma = torch.ones(K, device=dev); beta = torch.zeros(K, device=dev); bias = torch.zeros(N, device=dev)
ref = torch.nn.functional.layer_norm(x.double(), (K,), gamma.double(), beta.double(), 1e-6) @ W32.double().t()
# separate: LN kernel -> GEMM
y = torch.empty_like(x); mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
_abi.check(lib.nrv_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-6, y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), M, K, 0, sp))
o1 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
_abi.gemm(y, W, o1, bias=bias)
# folded
Wf = torch.empty_like(W); c = torch.empty(N, device=dev)
_abi.check(lib.nrv_ln_fold_weights(W.data_ptr(), gamma.data_ptr(), beta.data_ptr(), bias.data_ptr(), Wf.data_ptr(), c.data_ptr(), N, K, K, 0, sp))
stats = torch.empty(M, 2, device=dev, dtype=torch.float64)
_abi.check(lib.nrv_rowstats(x.data_ptr(), M, K, 0, stats.data_ptr(), sp))
o2 = torch.empty_like(o1)
_abi.gemm(x, Wf, o2, bias=c, ln_stats=stats, ln_eps=1e-6, K_ln=K)
o2f = torch.empty(M, N, device=dev)    # same product, fp32 output: isolates operand error from output rounding
_abi.gemm(x, Wf, o2f, bias=c, ln_stats=stats, ln_eps=1e-6, K_ln=K)
print("separate rel %.5f | folded rel %.5f | folded fp32-out rel %.5f" % (rel(o1, ref), rel(o2, ref), rel(o2f, ref)))
Wc = W.double() - W.double().mean(1, keepdim=True)
print("Wf vs centred bf16 W: rel %.5f ; row sums max %.2e ; changed elements per row: max |Wf - round(Wc)| %.2e" % (
    rel(Wf, Wc), Wf.double().sum(1).abs().max(), (Wf.double() - Wc.to(torch.bfloat16).double()).abs().max()))
xh = torch.nn.functional.layer_norm(x.double(), (K,), None, None, 1e-6)
print("exact LN x (Wf)^T vs ref: rel %.5f" % rel(xh @ Wf.double().t(), ref))
print("rs check: rel %.2e" % rel(1.0 / torch.sqrt(x.double().var(1, unbiased=False) + 1e-6), rstd))
