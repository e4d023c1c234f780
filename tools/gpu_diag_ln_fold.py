import os, sys, torch
sys.path.insert(0, "noise-robust-vit_b200"); sys.path.insert(0, "oracle")
import vit_pytorch_robust as V, vit_oracle as O
from vit_pytorch_robust import _abi
dev = "cuda:0"
for seed in (0, 1):
    torch.manual_seed(seed)
    m = V.vit_b_16()
    with torch.no_grad():
        m.heads.head.weight.normal_(std=0.02); m.class_token.normal_(std=0.02)
    m = m.to(dev).eval(); m._nrv.ln_fold_min_tokens = 0
    g = torch.Generator().manual_seed(5 + seed)
    img = torch.randn(64, 3, 224, 224, generator=g).to(torch.bfloat16).to(dev)
    outs = {}
    with torch.no_grad():
        for mode, tag in ((_abi.LN_FOLDED, "fold"), (_abi.LN_SEPARATE, "sep")):
            m._nrv.ln_mode_infer = mode; m._nrv._graphs.clear()
            outs[tag] = m(img).float()
        m._nrv.compute_dtype = torch.float32
        ref = m(img.float())
    print("seed %d: cos(fold, fp32) %.6f  cos(sep, fp32) %.6f  cos(fold, sep) %.6f | rel fold %.4f sep %.4f" % (
        seed, O.cosine(outs["fold"], ref), O.cosine(outs["sep"], ref), O.cosine(outs["fold"], outs["sep"]),
        O.rel_l2(outs["fold"], ref), O.rel_l2(outs["sep"], ref)))
    # per-token |mean| / std of the residual stream entering a few layers (fp32 path, hooks)
    seen = {}
    hs = [m.encoder.layers[l].register_forward_hook(lambda mod, i, o, l=l: seen.__setitem__(l, i[0])) for l in (0, 3, 6, 11)]
    with torch.no_grad():
        m(img[:8].float())
    for h in hs: h.remove()
    for l, x in seen.items():
        r = (x.mean(-1).abs() / x.std(-1))
        print("   layer %2d input: |mean|/std median %.3f max %.3f ; max |x| %.1f, std %.2f" % (l, r.median(), r.max(), x.abs().max(), x.std()))
