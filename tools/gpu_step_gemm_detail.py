"""Per-GEMM timing inside a real training step (CUDA events around every nrv_gemm launch)."""
import collections, ctypes as C, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
import vit_pytorch_robust as V
from vit_pytorch_robust import _abi
dev = torch.device("cuda:0")
B = 256
torch.manual_seed(0)
model = V.vit_b_16()
with torch.no_grad():
    model.heads.head.weight.normal_(std=0.02)
model = model.to(dev)
opt = V.FusedAdamW(model.parameters(), lr=2e-4, weight_decay=0.01)
img = torch.randn(B, 3, 224, 224, device=dev).to(torch.bfloat16)
lab = torch.randint(0, 1000, (B,), device=dev)
lib = _abi.load()
def step():
    opt.zero_grad()
    V.softmax_cross_entropy(model(img), lab, 0.1).backward()
    opt.step()
for _ in range(8):
    step()
torch.cuda.synchronize()
lib.nrv_gemm_timing(1)
for _ in range(3):
    step()
torch.cuda.synchronize()
buf = (C.c_longlong * (5 * 4096))()
n = lib.nrv_gemm_timing_detail(buf, 4096)
lib.nrv_gemm_timing(0)
agg = collections.OrderedDict()
for i in range(n):
    M, N, K, e, us = buf[5 * i:5 * i + 5]
    agg.setdefault((M, N, K, e), []).append(us)
tot = 0
names = {0: "store", 1: "gelu", 2: "dgelu", 3: "atomic", 4: "gelu_g", 5: "mul"}
for (M, N, K, e), v in agg.items():
    v.sort(); med = v[len(v) // 2]; tot += sum(v) / 3
    print("M=%6d N=%5d K=%6d epi=%-6s a_mn=%d b_mn=%d  n=%3d  median %7.1f us  %6.0f TF  (min %d max %d)" %
          (M, N, K, names[e & 15], (e >> 4) & 1, (e >> 5) & 1, len(v), med, 2.0 * M * N * K / med / 1e6, v[0], v[-1]))
print("GEMM total per step: %.2f ms" % (tot / 1e3))
