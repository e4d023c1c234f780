"""CUDA-event timing of nrv_layernorm_fwd / bwd and nrv_colsum at the ViT-B/16 shape (T = 50432 rows, dim 768)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
from vit_pytorch_robust import _abi
dev = torch.device("cuda:0"); lib = _abi.init(dev)
rows, dim = 50432, 768
x = torch.randn(rows, dim, device=dev).to(torch.bfloat16); dy = torch.randn_like(x); dres = torch.randn_like(x)
y = torch.empty_like(x); dx = torch.empty_like(x)
gamma = torch.ones(dim, device=dev); beta = torch.zeros(dim, device=dev)
mean = torch.empty(rows, device=dev); rstd = torch.empty(rows, device=dev)
dg = torch.zeros(dim, device=dev); db = torch.zeros(dim, device=dev); cs = torch.zeros(dim, device=dev)
nb = lib.nrv_layernorm_bwd_workspace(rows, dim); ws = torch.empty(nb, dtype=torch.uint8, device=dev)
sp = _abi.stream_ptr()
big = torch.randn(rows, 3072, device=dev).to(torch.bfloat16); bsum = torch.zeros(3072, device=dev)
nb2 = lib.nrv_colsum_workspace(rows, 3072); ws2 = torch.empty(nb2, dtype=torch.uint8, device=dev)
def fwd(): _abi.check(lib.nrv_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-6, y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, dim, 0, sp))
def bwd(): _abi.check(lib.nrv_layernorm_bwd(dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), dres.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), cs.data_ptr(), None, None, rows, dim, 0, ws.data_ptr(), nb, sp))
xn = torch.empty_like(x)
def bwd_xn(): _abi.check(lib.nrv_layernorm_bwd(dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), dres.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), cs.data_ptr(), beta.data_ptr(), xn.data_ptr(), rows, dim, 0, ws.data_ptr(), nb, sp))
stats = torch.empty(rows, 2, device=dev, dtype=torch.float64)
def rstats(): _abi.check(lib.nrv_rowstats(x.data_ptr(), rows, dim, 0, stats.data_ptr(), sp))
def csum(): _abi.check(lib.nrv_colsum(big.data_ptr(), 3072, rows, 3072, 0, bsum.data_ptr(), ws2.data_ptr(), nb2, sp))
for name, fn, nbytes in (("ln_fwd", fwd, rows * dim * 4), ("ln_bwd", bwd, rows * dim * 8), ("ln_bwd+xn", bwd_xn, rows * dim * 10), ("rowstats", rstats, rows * dim * 2), ("colsum 3072", csum, rows * 3072 * 2)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("%-12s %.1f us  %.2f TB/s (algorithmic bytes %d MB)" % (name, ms * 1e3, nbytes / ms / 1e9, nbytes >> 20), flush=True)
