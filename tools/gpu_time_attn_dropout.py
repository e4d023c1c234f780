"""Training step of ViT-B/16 (224x224, bf16) with attention_dropout = 0.1: general tcgen05 attention kernels drawing the
mask themselves against the CUDA-core kernels (attn_impl = SIMT) and against the step without dropout."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
import vit_pytorch_robust as V
from vit_pytorch_robust import _abi
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = torch.randn(B, 3, 224, 224, device=dev).to(torch.bfloat16)
y = torch.randint(0, 1000, (B,), device=dev)
for name, p_attn, impl in (("no dropout", 0.0, _abi.ATTN_IMPL_AUTO), ("attention_dropout 0.1, tcgen05 (general kernels)", 0.1, _abi.ATTN_IMPL_TC),
                           ("attention_dropout 0.1, CUDA cores", 0.1, _abi.ATTN_IMPL_SIMT)):
    torch.manual_seed(0)
    m = V.vit_b_16(attention_dropout=p_attn)
    with torch.no_grad():
        m.heads.head.weight.normal_(std=0.02)
    m = m.to(dev).train()
    m._nrv.attn_impl = impl
    opt = V.FusedAdamW(m.parameters(), lr=2e-4, weight_decay=0.01)
    def step():
        opt.zero_grad()
        loss = V.softmax_cross_entropy(m(x), y, 0.1)
        loss.backward()
        opt.step()
        return loss
    for _ in range(3):
        out = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("ViT-B/16 B=%d %-50s %8.2f ms/step %7.0f img/s (loss %.4f)" % (B, name, ms, B / ms * 1e3, float(out.detach())), flush=True)
    del m, opt
