"""Critical-path breakdown of the tcgen05 attention kernels (control-thread clock64 stamps, CTA 0)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
from vit_pytorch_robust import _abi
dev = torch.device("cuda:0"); lib = _abi.init(dev)
B, N, H, dh = 256, 197, 12, 64
qkv = torch.randn(B, N, 3 * H * dh, device=dev).to(torch.bfloat16)
dout = torch.randn(B, N, H * dh, device=dev).to(torch.bfloat16)
out = torch.empty_like(dout); dqkv = torch.empty_like(qkv); lse = torch.empty(B, H, N, device=dev)
nb = lib.nrv_attn_bwd_workspace(B, N, H, dh); ws = torch.empty(nb, dtype=torch.uint8, device=dev)
sp = _abi.stream_ptr()
names = ["start->rows", "rows->cols", "cols->mma1 issued", "mma1->done(bar_s)", "bar_s->softmax done(bar_p)", "->PV issued", "->epilogue done"]
def show(tag, buf):
    t = buf.cpu()[:512].view(64, 8)[:6]
    u = buf.cpu()[512:].view(64, 8)[:6]
    print(tag)
    for g in range(6):
        d = [(t[g, i + 1] - t[g, i]).item() for i in range(7)]
        print("  tile %2d total %6d : " % (g, (t[g, 7] - t[g, 0]).item()) + "  ".join("%s=%d" % (n, v) for n, v in zip(names, d)))
        print("     softmax warp0: tmem_ld=%d max+xchg=%d exp+store=%d fence+arrive=%d wait_o=%d epilogue=%d" % tuple((u[g, i + 1] - u[g, i]).item() for i in range(6)))
for mode in ("fwd", "bwd"):
    buf = torch.zeros(2 * 64 * 8, dtype=torch.int64, device=dev)
    for it in range(2):
        lib.nrv_attn_debug_timestamps(buf.data_ptr())
        if mode == "fwd":
            _abi.check(lib.nrv_attn_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, dh, dh ** -0.5, 0, 0, 2, None, 0, sp))
        else:
            _abi.check(lib.nrv_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), B, N, H, dh, dh ** -0.5, 0, 0, 2, ws.data_ptr(), nb, sp))
        torch.cuda.synchronize()
    lib.nrv_attn_debug_timestamps(None)
    show(mode + " (bwd shows the DKV pass, which overwrites the DQ stamps)", buf)
