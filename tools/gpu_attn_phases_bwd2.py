"""Phase timestamps of the fused attention backward (CTA 0, first 40 blocks)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
from vit_pytorch_robust import _abi
dev = torch.device("cuda:0"); lib = _abi.init(dev)
B, N, H, dh = 256, 197, 12, 64
qkv = torch.randn(B, N, 3 * H * dh, device=dev).to(torch.bfloat16)
dout = torch.randn(B, N, H * dh, device=dev).to(torch.bfloat16)
out = torch.empty_like(dout); dqkv = torch.empty_like(qkv); lse = torch.zeros(B, H, N, device=dev)
nb = lib.nrv_attn_bwd_workspace(B, N, H, dh); ws = torch.empty(nb, dtype=torch.uint8, device=dev)
sp = _abi.stream_ptr()
_abi.check(lib.nrv_attn_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, dh, dh ** -0.5, 0, 0, 2, None, 0, sp))
buf = torch.zeros(8 * 64, dtype=torch.int64, device=dev)
for it in range(2):
    lib.nrv_attn_debug_timestamps(buf.data_ptr())
    _abi.check(lib.nrv_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), B, N, H, dh, dh ** -0.5, 0, 0, 2, ws.data_ptr(), nb, sp))
    torch.cuda.synchronize()
lib.nrv_attn_debug_timestamps(None)
t = buf.cpu().view(64, 8)
t0 = t[0, 4].item()
for g in range(40):
    c = t[g] - t0
    print("blk %2d (item %d lb %d) mma: P-seen %6d acc-free %6d issued %6d | simt: S-seen %6d done %6d after-epi %6d" %
          (g, g // 8, g % 8, c[0], c[1], c[2], c[4], c[5], c[6]))
