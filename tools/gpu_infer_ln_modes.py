"""Inference forward (eval + no_grad, bf16) with the LayerNorms folded into the QKV / FC1 GEMMs vs stand-alone kernels."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
import vit_pytorch_robust as V
from vit_pytorch_robust import _abi
dev = torch.device("cuda:0")
def t(fn, n=15):
    for _ in range(4): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, ctor, B in (("vit_b_16", V.vit_b_16, 256), ("vit_b_16", V.vit_b_16, 8), ("vit_l_16", V.vit_l_16, 128), ("vit_h_14", V.vit_h_14, 64)):
    torch.manual_seed(0)
    m = ctor().to(dev).eval()
    m._nrv.ln_fold_min_tokens = 0
    x = torch.randn(B, 3, 224, 224, device=dev).to(torch.bfloat16)
    res = {}
    for mode, tag in ((_abi.LN_SEPARATE, "separate"), (_abi.LN_FOLDED, "folded")):
        m._nrv.ln_mode_infer = mode
        m._nrv._graphs.clear()
        with torch.no_grad():
            ms = t(lambda: m(x))
        res[tag] = ms
    print("%s B=%d  separate %.3f ms (%.0f img/s) | folded %.3f ms (%.0f img/s) | %+.1f %%" %
          (name, B, res["separate"], B / res["separate"] * 1e3, res["folded"], B / res["folded"] * 1e3,
           100 * (res["separate"] / res["folded"] - 1)), flush=True)
    del m
