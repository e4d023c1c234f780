"""A/B of the two tilings of the CTA-pair GEMM (tile_mode 1: 256x256 units, double-buffered accumulator; tile_mode 2:
512x256 units, B shared by two row blocks) on the ViT-B/16 B=256 step shapes.  CUDA events, 20 back-to-back launches
after 5 warm-up launches; inputs (78-310 MB) exceed nothing but are re-read from L2/HBM as in the step."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
from vit_pytorch_robust import _abi
dev = torch.device("cuda:0"); _abi.init(dev)
T, D, M3, Q = 50432, 768, 3072, 2304
def t(fn, n=20):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
def rnd(*s): return (torch.randn(*s, device=dev) / 8).to(torch.bfloat16)
cases = []
# (name, M, N, K, a, a_layout, b, b_layout, kwargs)
x768, x3072, x2304 = rnd(T, D), rnd(T, M3), rnd(T, Q)
w_fc1, w_fc2, w_qkv, w_out = rnd(M3, D), rnd(D, M3), rnd(Q, D), rnd(D, D)
bias768 = torch.zeros(D, device=dev); res = rnd(T, D)
o768, o3072, o2304 = torch.empty(T, D, device=dev, dtype=torch.bfloat16), torch.empty(T, M3, device=dev, dtype=torch.bfloat16), torch.empty(T, Q, device=dev, dtype=torch.bfloat16)
g32 = lambda m, n: torch.zeros(m, n, device=dev)
K_, MN = _abi.NRV_K_MAJOR, _abi.NRV_MN_MAJOR
cases = [
 ("fc2 fwd  T x768 x3072 +bias+res", x3072, K_, w_fc2, K_, o768, dict(bias=bias768, residual=res)),
 ("fc1 dX   T x768 x3072 (W mn)   ", x3072, K_, w_fc1, MN, o768, {}),
 ("qkv dX   T x768 x2304 (W mn)   ", x2304, K_, w_qkv, MN, o768, {}),
 ("qkv fwd  T x2304x768  +bias    ", x768, K_, w_qkv, K_, o2304, dict(bias=torch.zeros(Q, device=dev))),
 ("fc1 plain T x3072x768          ", x768, K_, w_fc1, K_, o3072, {}),
 ("out fwd  T x768 x768 +bias+res ", x768, K_, w_out, K_, o768, dict(bias=bias768, residual=res)),
 ("dW fc1   3072x768 xT  atomic   ", x3072, MN, x768, MN, g32(M3, D), dict(epi=_abi.EPI_ATOMIC_F32)),
 ("dW fc2   768 x3072xT  atomic   ", x768, MN, x3072, MN, g32(D, M3), dict(epi=_abi.EPI_ATOMIC_F32)),
 ("dW qkv   2304x768 xT  atomic   ", x2304, MN, x768, MN, g32(Q, D), dict(epi=_abi.EPI_ATOMIC_F32)),
 ("dW out   768 x768 xT  atomic   ", x768, MN, x768, MN, g32(D, D), dict(epi=_abi.EPI_ATOMIC_F32)),
]
for name, a, al, b, bl, out, kw in cases:
    Mm = a.shape[0] if al == K_ else a.shape[1]
    Kk = a.shape[1] if al == K_ else a.shape[0]
    Nn = b.shape[0] if bl == K_ else b.shape[1]
    r = []
    for mode in (1, 2):
        us = t(lambda: _abi.gemm(a, b, out, a_layout=al, b_layout=bl, tile_mode=mode, **kw))
        r.append((us, 2.0 * Mm * Nn * Kk / us / 1e6))
    print("%s  classic %7.1f us %5.0f TF | dual %7.1f us %5.0f TF | %+5.1f %%" % (name, r[0][0], r[0][1], r[1][0], r[1][1], 100 * (r[0][0] / r[1][0] - 1)))
