"""Phase timestamps of the second-generation attention kernels (CTA 0, both groups)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
from vit_pytorch_robust import _abi
dev = torch.device("cuda:0"); lib = _abi.init(dev)
B, N, H, dh = 256, 197, 12, 64
qkv = torch.randn(B, N, 3 * H * dh, device=dev).to(torch.bfloat16)
out = torch.empty(B, N, H * dh, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, H, N, device=dev)
sp = _abi.stream_ptr()
buf = torch.zeros(2 * 32 * 16, dtype=torch.int64, device=dev)
for it in range(2):
    lib.nrv_attn_debug_timestamps(buf.data_ptr())
    _abi.check(lib.nrv_attn_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, dh, dh ** -0.5, 0, 0, 2, None, 0, sp))
    torch.cuda.synchronize()
lib.nrv_attn_debug_timestamps(None)
t = buf.cpu().view(2, 32, 16)
t0 = t[0, 0, 0].item()
for g in range(2):
    print("group", g)
    for gi in range(8):
        c = [(t[g, gi, i].item() - t0) for i in range(5)]
        s = [(t[g, gi, 8 + i].item() - t0) for i in range(7)]
        print("  tile %d ctrl: start %6d  S-issue %6d  S-done %6d  P-ready %6d  PV-issued %6d" % (gi, *c))
        print("         smx : begin %6d  S-seen %6d  max-done %6d  exp-done %6d  (arrive) %6d  O-seen %6d  epi-done %6d" % tuple(s))
