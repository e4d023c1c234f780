"""The four forward GEMMs of a ViT-B/16 layer (B=256) with and without the folded-LayerNorm epilogue variants."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
from vit_pytorch_robust import _abi
dev = torch.device("cuda:0"); lib = _abi.init(dev)
T, D, M3, Q = 50432, 768, 3072, 2304
def t(fn, n=20):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
def rnd(*s): return (torch.randn(*s, device=dev) / 8).to(torch.bfloat16)
x, h3072 = rnd(T, D), rnd(T, M3)
w_qkv, w_fc1, w_out, w_fc2 = rnd(Q, D), rnd(M3, D), rnd(D, D), rnd(D, M3)
bq, b1, bo = torch.zeros(Q, device=dev), torch.zeros(M3, device=dev), torch.zeros(D, device=dev)
stats = torch.zeros(T, 2, device=dev, dtype=torch.float64); stats[:, 1] = D
st_out = torch.zeros(T, 2, device=dev, dtype=torch.float64)
mean, rstd = torch.empty(T, device=dev), torch.empty(T, device=dev)
oq, o1, g1, od = (torch.empty(T, Q, device=dev, dtype=torch.bfloat16), torch.empty(T, M3, device=dev, dtype=torch.bfloat16),
                  torch.empty(T, M3, device=dev, dtype=torch.bfloat16), torch.empty(T, D, device=dev, dtype=torch.bfloat16))
res = rnd(T, D)
ln = dict(ln_stats=stats, ln_eps=1e-6, K_ln=D, ln_mean_out=mean, ln_rstd_out=rstd)
print("qkv   plain %.1f | folded LN %.1f us" % (t(lambda: _abi.gemm(x, w_qkv, oq, bias=bq)), t(lambda: _abi.gemm(x, w_qkv, oq, bias=bq, **ln))))
print("fc1   gelu+g %.1f | folded LN %.1f us" % (t(lambda: _abi.gemm(x, w_fc1, o1, bias=b1, epi=_abi.EPI_GELU_GRAD, out2=g1)),
                                                t(lambda: _abi.gemm(x, w_fc1, o1, bias=b1, epi=_abi.EPI_GELU_GRAD, out2=g1, **ln))))
print("fc1   gelu   %.1f | folded LN %.1f us" % (t(lambda: _abi.gemm(x, w_fc1, o1, bias=b1, epi=_abi.EPI_GELU)),
                                                t(lambda: _abi.gemm(x, w_fc1, o1, bias=b1, epi=_abi.EPI_GELU, **ln))))
print("out   +res  %.1f | + stats_out %.1f us" % (t(lambda: _abi.gemm(x, w_out, od, bias=bo, residual=res)), t(lambda: _abi.gemm(x, w_out, od, bias=bo, residual=res, stats_out=st_out))))
print("fc2   +res  %.1f | + stats_out %.1f us" % (t(lambda: _abi.gemm(h3072, w_fc2, od, bias=bo, residual=res)), t(lambda: _abi.gemm(h3072, w_fc2, od, bias=bo, residual=res, stats_out=st_out))))
Wf = torch.empty_like(w_qkv); gam = torch.ones(D, device=dev); bet = torch.zeros(D, device=dev); cq = torch.empty(Q, device=dev)
print("fold weights (one 2304x768 matrix) %.1f us" % t(lambda: _abi.check(lib.nrv_ln_fold_weights(w_qkv.data_ptr(), gam.data_ptr(), bet.data_ptr(), bq.data_ptr(), Wf.data_ptr(), cq.data_ptr(), Q, D, D, 0, _abi.stream_ptr()))))
