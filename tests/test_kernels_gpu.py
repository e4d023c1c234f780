"""Per-kernel parity on the B200: every libnrvit entry against a plain torch fp32/fp64 statement of
the same op on the same seeded inputs, called through the C ABI (ctypes)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

from vit_pytorch_robust import _abi  # noqa: E402


def dev():
    return torch.device("cuda:0")


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def tol(dtype, f32=2e-5, bf16=6e-3):
    return f32 if dtype == torch.float32 else bf16


def sp():
    return _abi.stream_ptr()


DT = [torch.bfloat16, torch.float32]


# ---------------------------------------------------------------- GEMM
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("shape", [(128, 256, 64), (200, 136, 72), (776, 520, 328), (64, 1000, 768), (2056, 776, 200)])
@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("single", [0, 1], ids=["pair", "single"])
def test_gemm_layouts(a_mn, b_mn, shape, dtype, single):
    M, N, K = shape
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K)
    A = torch.randn(M, K, generator=g).to(dev(), dtype)
    B = torch.randn(N, K, generator=g).to(dev(), dtype)
    out = torch.full((M, N), float("nan"), device=dev(), dtype=dtype)
    _abi.gemm(A.t().contiguous() if a_mn else A, B.t().contiguous() if b_mn else B, out,
              a_layout=a_mn, b_layout=b_mn, M=M, N=N, K=K, force_single_cta=single)
    ref = A.double() @ B.double().t()
    assert rel(out, ref) < tol(dtype)


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("shape", [(512, 256, 64), (1024, 512, 1600), (776, 520, 328), (1288, 264, 2304), (2056, 776, 200),
                                   (5000, 768, 3072),
                                   # odd row-block counts with more units than CTA pairs: the half-dead units of the last row are
                                   # dealt to the pairs that carry an extra unit (unit_decode): 41 / 197 / 151 row blocks
                                   (10400, 1000, 136), (50432, 768, 192), (38504, 264, 72)])
def test_gemm_dual_accumulator_tiles(a_mn, b_mn, shape):
    """tile_mode=2: 512x256 work units (two row blocks share the B tile, both TMEM accumulators live).  Ragged M (a dead
    or partial second row block), ragged N and K, every operand layout; plain store, bias + residual, split-K atomics."""
    M, N, K = shape
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K)
    A = torch.randn(M, K, generator=g).to(dev(), torch.bfloat16)
    B = (torch.randn(N, K, generator=g) / 8).to(dev(), torch.bfloat16)
    a_op = A.t().contiguous() if a_mn else A
    b_op = B.t().contiguous() if b_mn else B
    ref = A.double() @ B.double().t()
    kw = dict(a_layout=a_mn, b_layout=b_mn, M=M, N=N, K=K, tile_mode=2)
    out = torch.full((M, N), float("nan"), device=dev(), dtype=torch.bfloat16)
    _abi.gemm(a_op, b_op, out, **kw)
    assert rel(out, ref) < 6e-3
    bias = torch.randn(N, generator=g).to(dev())
    res = torch.randn(M, N, generator=g).to(dev(), torch.bfloat16)
    out.fill_(float("nan"))
    _abi.gemm(a_op, b_op, out, bias=bias, residual=res, **kw)
    assert rel(out, ref + bias.double() + res.double()) < 6e-3
    acc = torch.ones(M, N, device=dev(), dtype=torch.float32)
    _abi.gemm(a_op, b_op, acc, epi=_abi.EPI_ATOMIC_F32, **kw)
    assert rel(acc, ref + 1.0) < 6e-3
    # the classic tiling of the same product agrees to accumulation-order noise
    out1 = torch.empty_like(out)
    _abi.gemm(a_op, b_op, out1, bias=bias, residual=res, a_layout=a_mn, b_layout=b_mn, M=M, N=N, K=K, tile_mode=1)
    assert rel(out, out1) < 2e-3


@pytest.mark.parametrize("dtype", DT)
def test_gemm_epilogues(dtype):
    M, N, K = 384, 520, 256
    g = torch.Generator().manual_seed(5)
    A = torch.randn(M, K, generator=g).to(dev(), dtype)
    B = (torch.randn(N, K, generator=g) / 16).to(dev(), dtype)
    bias = torch.randn(N, generator=g).to(dev())
    res = torch.randn(M, N, generator=g).to(dev(), dtype)
    acc = A.double() @ B.double().t()
    t = tol(dtype)
    out = torch.empty(M, N, device=dev(), dtype=dtype)
    _abi.gemm(A, B, out, bias=bias, residual=res, alpha=0.5)
    assert rel(out, 0.5 * acc + bias.double() + res.double()) < t
    out2 = torch.empty_like(out)
    _abi.gemm(A, B, out, bias=bias, epi=_abi.EPI_GELU, out2=out2)
    u = acc + bias.double()
    assert rel(out, torch.nn.functional.gelu(u)) < t
    assert rel(out2, u) < t
    aux = torch.randn(M, N, generator=g).to(dev(), dtype)
    _abi.gemm(A, B, out, epi=_abi.EPI_DGELU, aux=aux)
    x = aux.double().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    assert rel(out, acc * x.grad) < t
    # training pair: FC1 epilogue stores gelu(u) and gelu'(u); the FC2 dX epilogue multiplies by the stored factor
    g2 = torch.empty_like(out)
    _abi.gemm(A, B, out, bias=bias, epi=_abi.EPI_GELU_GRAD, out2=g2)
    xu = u.clone().requires_grad_(True)
    torch.nn.functional.gelu(xu).sum().backward()
    assert rel(out, torch.nn.functional.gelu(u)) < t
    assert rel(g2, xu.grad) < t
    # absolute accuracy of the sigmoid-form GELU of the bf16 path (fit: 5e-6 / 1.8e-5) is far below bf16 rounding
    if dtype == torch.bfloat16:
        assert (out.double() - torch.nn.functional.gelu(u)).abs().max() < 4e-2
        assert (g2.double() - xu.grad).abs().max() < 8e-3
    _abi.gemm(A, B, out, epi=_abi.EPI_MUL, aux=aux)
    assert rel(out, acc * aux.double()) < t
    if dtype == torch.bfloat16:
        # fused bias gradient: colsum += column sums of the tile as stored (bf16), accumulated in fp32
        cs = torch.full((N,), 0.5, device=dev(), dtype=torch.float32)
        _abi.gemm(A, B, out, epi=_abi.EPI_MUL, aux=aux, colsum=cs)
        assert rel(out, acc * aux.double()) < t
        assert rel(cs, 0.5 + out.double().sum(0)) < 1e-5
    else:
        with pytest.raises(_abi.NrvError):
            _abi.gemm(A, B, out, epi=_abi.EPI_MUL, aux=aux, colsum=torch.zeros(N, device=dev()))
    with pytest.raises(_abi.NrvError):
        _abi.gemm(A, B, out, epi=_abi.EPI_GELU_GRAD)          # needs out2
    with pytest.raises(_abi.NrvError):
        _abi.gemm(A, B, out, epi=_abi.EPI_MUL)                # needs aux
    outf = torch.zeros(M, N, device=dev(), dtype=torch.float32)
    _abi.gemm(A, B, outf, epi=_abi.EPI_ATOMIC_F32, splits=3)
    _abi.gemm(A, B, outf, epi=_abi.EPI_ATOMIC_F32)
    assert rel(outf, 2 * acc) < 2e-5
    # token-row remap + positional table (patch embedding with a class token)
    Bsz, n_in, n_out = 6, 64, 65
    pos = torch.randn(n_out, N, generator=g).to(dev())
    outp = torch.zeros(Bsz * n_out, N, device=dev(), dtype=dtype)
    _abi.gemm(A, B, outp, bias=bias, pos=pos, pos_rows_in=n_in, pos_rows_out=n_out, pos_row_off=1)
    refp = torch.zeros(Bsz, n_out, N, dtype=torch.double, device=dev())
    refp[:, 1:] = (acc + bias.double()).view(Bsz, n_in, N) + pos.double()[1:]
    assert rel(outp.view(Bsz, n_out, N), refp) < t


def test_gelu_epilogue_error_is_bf16_rounding():
    """The bf16 epilogues evaluate GELU / GELU' as Phi = 0.5 + 0.5 tanh.approx(u P(u^2)) (common.cuh: gelu_sig_pair).  Against the
    exact erf forms on fp64 the stored bf16 values must be as close as the exactly computed values rounded to bf16 -- over all
    elements and separately in the negative tail, where Phi is small and an absolute error of the tanh would show."""
    M, N, K = 2048, 1024, 256
    g = torch.Generator().manual_seed(11)
    A = torch.randn(M, K, generator=g).to(dev(), torch.bfloat16)
    B = (torch.randn(N, K, generator=g) / 8).to(dev(), torch.bfloat16)       # u ~ N(-0.5, 2^2): plenty of |u| in 1.5 .. 6
    bias = (torch.randn(N, generator=g) * 0.5 - 0.5).to(dev())
    h = torch.empty(M, N, device=dev(), dtype=torch.bfloat16)
    gp = torch.empty_like(h)
    _abi.gemm(A, B, h, bias=bias, epi=_abi.EPI_GELU_GRAD, out2=gp)
    h_inf = torch.empty_like(h)
    _abi.gemm(A, B, h_inf, bias=bias, epi=_abi.EPI_GELU)
    u = (A.double() @ B.double().t() + bias.double()).requires_grad_(True)
    ref_h = torch.nn.functional.gelu(u)
    ref_h.sum().backward()
    ref_g, ref_h, u = u.grad, ref_h.detach(), u.detach()

    def rms(x):
        return x.pow(2).mean().sqrt().item()
    for got, ref in ((h, ref_h), (h_inf, ref_h), (gp, ref_g)):
        for sel in (torch.ones_like(u, dtype=torch.bool), u < -1.5, u > 1.5):
            err = rms((got.double() - ref)[sel])
            rounding = rms((ref.to(torch.bfloat16).double() - ref)[sel])
            assert err < 1.15 * rounding + 1e-7, (err, rounding)
    assert torch.equal(h, h_inf)


def test_gemm_rejects_bad_arguments():
    A = torch.zeros(16, 64, device=dev(), dtype=torch.bfloat16)
    out = torch.zeros(16, 12, device=dev(), dtype=torch.bfloat16)
    with pytest.raises(_abi.NrvError):
        _abi.gemm(A, torch.zeros(12, 64, device=dev(), dtype=torch.bfloat16), out)  # N % 8 != 0
    with pytest.raises(_abi.NrvError):
        _abi.gemm(A.cpu(), A.cpu(), out.cpu())


# ---------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("rows,dim", [(1, 64), (37, 512), (1000, 768), (513, 1024), (64, 1280)])
def test_layernorm_fwd_bwd(dtype, rows, dim):
    lib = _abi.init(dev())
    g = torch.Generator().manual_seed(rows + dim)
    x = (torch.randn(rows, dim, generator=g) * 2 + 0.5).to(dev(), dtype)
    gamma = (1 + 0.1 * torch.randn(dim, generator=g)).to(dev())
    beta = (0.1 * torch.randn(dim, generator=g)).to(dev())
    dy = torch.randn(rows, dim, generator=g).to(dev(), dtype)
    dres = torch.randn(rows, dim, generator=g).to(dev(), dtype)
    y = torch.empty_like(x)
    mean = torch.empty(rows, device=dev())
    rstd = torch.empty(rows, device=dev())
    code = _abi._dt(x)
    _abi.check(lib.nrv_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-5, y.data_ptr(),
                                     mean.data_ptr(), rstd.data_ptr(), rows, dim, code, sp()))
    xd = x.double().requires_grad_(True)
    gd = gamma.double().requires_grad_(True)
    bd = beta.double().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xd, (dim,), gd, bd, 1e-5)
    assert rel(y, yr) < tol(dtype, 2e-6)
    assert rel(mean, xd.mean(-1)) < 1e-5
    yr.backward(dy.double())
    dx = torch.empty_like(x)
    dgam = torch.ones(dim, device=dev())
    dbet = torch.ones(dim, device=dev())
    csum = torch.zeros(dim, device=dev())
    nb = lib.nrv_layernorm_bwd_workspace(rows, dim)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev())
    _abi.check(lib.nrv_layernorm_bwd(dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                     dres.data_ptr(), dx.data_ptr(), dgam.data_ptr(), dbet.data_ptr(), csum.data_ptr(),
                                     None, None, rows, dim, code, ws.data_ptr(), nb, sp()))
    want_dx = xd.grad + dres.double()
    assert rel(dx, want_dx) < tol(dtype, 2e-5)
    assert rel(dgam - 1, gd.grad) < tol(dtype, 2e-5, 2e-3)   # accumulates onto the existing value
    assert rel(dbet - 1, bd.grad) < tol(dtype, 2e-5, 2e-3)
    assert rel(csum, dx.double().sum(0)) < 1e-4
    # same call, also asked for the normalised rows (forward passes with the LayerNorm folded into the GEMM do not keep them)
    dx2 = torch.empty_like(x)
    xn = torch.full_like(x, float("nan"))
    dgam2, dbet2, csum2 = torch.ones(dim, device=dev()), torch.ones(dim, device=dev()), torch.zeros(dim, device=dev())
    _abi.check(lib.nrv_layernorm_bwd(dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                     dres.data_ptr(), dx2.data_ptr(), dgam2.data_ptr(), dbet2.data_ptr(), csum2.data_ptr(),
                                     beta.data_ptr(), xn.data_ptr(), rows, dim, code, ws.data_ptr(), nb, sp()))
    assert torch.equal(dx2, dx)
    assert rel(xn, yr.detach()) < tol(dtype, 2e-6)
    assert rel(dgam2 - 1, gd.grad) < tol(dtype, 2e-5, 2e-3) and rel(csum2, dx.double().sum(0)) < 1e-4


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("M,N,K,offset", [(300, 264, 192, 0.0), (1000, 776, 768, 0.5), (517, 2304, 768, 10.0), (4100, 3072, 1024, -3.0)])
def test_layernorm_folded_into_the_gemm(dtype, M, N, K, offset):
    """The fused LayerNorm + projection of the forward pass (simple_vit.py:65-67,38-39 ; vit.py:123,128):
    GEMM on the raw rows with B = gamma o W and the two row statistics applied in the epilogue, against
    LayerNorm -> Linear on fp64.  offset: common shift of every channel in units of the channel std (the E[x^2] - mu^2 form
    loses accuracy only for |mu| >> sigma).  Also the GELU + GELU' epilogue and the statistics written by a producing GEMM."""
    lib = _abi.init(dev())
    g = torch.Generator().manual_seed(M + N + K)
    x = (torch.randn(M, K, generator=g) * 1.5 + 1.5 * offset).to(dev(), dtype)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev(), dtype)
    gamma = (1 + 0.2 * torch.randn(K, generator=g)).to(dev())
    beta = (0.2 * torch.randn(K, generator=g)).to(dev())
    bias = torch.randn(N, generator=g).to(dev())
    code = _abi._dt(x)
    Wf = torch.empty_like(W)
    c_vec = torch.empty(N, device=dev())
    _abi.check(lib.nrv_ln_fold_weights(W.data_ptr(), gamma.data_ptr(), beta.data_ptr(), bias.data_ptr(), Wf.data_ptr(),
                                       c_vec.data_ptr(), N, K, K, code, sp()))
    gw = W.double() * gamma.double()
    assert rel(Wf, gw - gw.mean(1, keepdim=True)) < tol(dtype, 1e-6, 8e-3)
    assert rel(c_vec, W.double() @ beta.double() + bias.double()) < 1e-5
    # the stored rows sum to zero far below their own rounding noise: the mean of x cancels inside the accumulation
    row_scale = Wf.double().abs().sum(1)
    assert (Wf.double().sum(1).abs() / row_scale).max() < (2e-7 if dtype == torch.float32 else 2e-5)
    stats = torch.full((M, 2), float("nan"), device=dev(), dtype=torch.float64)
    _abi.check(lib.nrv_rowstats(x.data_ptr(), M, K, code, stats.data_ptr(), sp()))
    assert rel(stats[:, 0], x.double().sum(1)) < 1e-5 and rel(stats[:, 1], x.double().pow(2).sum(1)) < 1e-5
    out = torch.full((M, N), float("nan"), device=dev(), dtype=dtype)
    mean, rstd = torch.empty(M, device=dev()), torch.empty(M, device=dev())
    _abi.gemm(x, Wf, out, bias=c_vec, ln_stats=stats, ln_eps=1e-6, K_ln=K, ln_mean_out=mean, ln_rstd_out=rstd)
    xd = x.double()
    ref_n = torch.nn.functional.layer_norm(xd, (K,), gamma.double(), beta.double(), 1e-6)
    ref = ref_n @ W.double().t() + bias.double()
    # bf16: the operand roundings differ from LayerNorm -> bf16 -> GEMM (x stays exact, gamma o W is rounded), same error class
    assert rel(out, ref) < tol(dtype, 3e-5 if abs(offset) <= 1 else 2e-4, 8e-3 if abs(offset) <= 1 else 1.2e-2)
    assert rel(mean, xd.mean(1)) < 1e-5
    assert rel(rstd, 1.0 / torch.sqrt(xd.var(1, unbiased=False) + 1e-6)) < (1e-5 if abs(offset) <= 1 else 1e-3)
    if abs(offset) <= 1:
        h, gp = torch.empty_like(out), torch.empty_like(out)
        _abi.gemm(x, Wf, h, bias=c_vec, ln_stats=stats, ln_eps=1e-6, K_ln=K, epi=_abi.EPI_GELU_GRAD, out2=gp)
        u = ref.clone().requires_grad_(True)
        torch.nn.functional.gelu(u).sum().backward()
        assert rel(h, torch.nn.functional.gelu(ref)) < tol(dtype, 5e-5, 1e-2) and rel(gp, u.grad) < tol(dtype, 5e-5, 1e-2)
    # statistics emitted by a producing GEMM (bias + residual epilogue): row sums of its output (before the store rounds it)
    res = torch.randn(M, N, generator=g).to(dev(), dtype)
    st2 = torch.zeros(M, 2, device=dev(), dtype=torch.float64)
    y = torch.empty(M, N, device=dev(), dtype=dtype)
    _abi.gemm(x, W, y, bias=bias, residual=res, stats_out=st2)
    yd = x.double() @ W.double().t() + bias.double() + res.double()
    t1 = 5e-5 if dtype == torch.float32 else 2e-3          # the fp32 accumulator against fp64; mean of N values with noise
    assert (st2[:, 0] - yd.sum(1)).abs().max() < t1 * yd.abs().sum(1).max() and rel(st2[:, 1], yd.pow(2).sum(1)) < t1
    st3 = torch.zeros(M, 2, device=dev(), dtype=torch.float64)
    _abi.gemm(x, W, y, stats_out=st3)                      # plain store
    yd = x.double() @ W.double().t()
    assert (st3[:, 0] - yd.sum(1)).abs().max() < t1 * yd.abs().sum(1).max() and rel(st3[:, 1], yd.pow(2).sum(1)) < t1


@pytest.mark.parametrize("dtype", DT)
def test_colsum(dtype):
    lib = _abi.init(dev())
    x = torch.randn(3001, 776, device=dev()).to(dtype)
    out = torch.full((776,), 2.0, device=dev())
    nb = lib.nrv_colsum_workspace(3001, 776)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev())
    _abi.check(lib.nrv_colsum(x.data_ptr(), 776, 3001, 776, _abi._dt(x), out.data_ptr(), ws.data_ptr(), nb, sp()))
    assert rel(out - 2.0, x.double().sum(0)) < 1e-4


# ---------------------------------------------------------------- patches / tokens / pooling
@pytest.mark.parametrize("order", [_abi.PATCH_P1P2C, _abi.PATCH_CP1P2])
@pytest.mark.parametrize("dtype", DT)
def test_im2col(order, dtype):
    import vit_oracle as O
    lib = _abi.init(dev())
    B, Cc, H, W, ph, pw = 3, 3, 28, 42, 14, 7
    img = torch.randn(B, Cc, H, W, device=dev())
    kdim = Cc * ph * pw
    ld = (kdim + 7) // 8 * 8
    n = (H // ph) * (W // pw)
    out = torch.full((B * n, ld), float("nan"), device=dev(), dtype=dtype)
    _abi.check(lib.nrv_im2col(img.data_ptr(), _abi.NRV_F32, B, Cc, H, W, ph, pw, order, out.data_ptr(), _abi._dt(out),
                              ld, sp()))
    fn = O.patchify_p1p2c if order == _abi.PATCH_P1P2C else O.patchify_cp1p2
    ref = fn(img.cpu(), ph, pw).reshape(B * n, kdim)
    assert torch.equal(out[:, :kdim].float().cpu(), ref.to(dtype).float())
    assert (out[:, kdim:] == 0).all()


@pytest.mark.parametrize("B,Cc,H,W,ph,pw,D,cls", [
    (5, 3, 224, 224, 16, 16, 768, 1),      # ViT-B/16: 14 patches per row -> tiles of 2 patches x 64 images, mostly out of bounds
    (70, 3, 64, 96, 16, 16, 256, 0),       # two image groups, the second one partial
    (8, 3, 256, 256, 32, 32, 1024, 1),     # README config: 8 patches per row -> tiles of 8 patches x 16 images
    (3, 3, 128, 192, 64, 64, 192, 1),      # 64-wide patches: one K block = one patch row (the 128-byte-swizzle case)
    (9, 4, 64, 64, 32, 16, 136, 0),        # rectangular patches, ragged D
    (130, 3, 48, 112, 16, 16, 264, 1),     # 7 patches per row (odd): tiles of 1 patch x 128 images, 2 image groups
])
def test_patch_embedding_with_im2col_fused_into_the_gemm(B, Cc, H, W, ph, pw, D, cls):
    """nrv_patch_embed_fwd / _bwd_weight (the im2col happens inside the GEMM's TMA loads) against the oracle's patchify +
    fp64 matmul (conv_proj order, vit.py:323-331), including the class-token row offset and the positional table."""
    import vit_oracle as O
    lib = _abi.init(dev())
    bf = _abi.NRV_BF16
    assert lib.nrv_patch_embed_supported(Cc, H, W, ph, pw, _abi.PATCH_CP1P2, bf, bf, D) == 1
    g = torch.Generator().manual_seed(B * 7 + D)
    n = (H // ph) * (W // pw)
    N = n + cls
    kdim = Cc * ph * pw
    img = torch.randn(B, Cc, H, W, generator=g).to(dev(), torch.bfloat16)
    w = (torch.randn(D, kdim, generator=g) / kdim ** 0.5).to(dev(), torch.bfloat16)
    bias = torch.randn(D, generator=g).to(dev())
    pos = torch.randn(N, D, generator=g).to(dev())
    out = torch.full((B * N, D), 7.0, device=dev(), dtype=torch.bfloat16)
    _abi.check(lib.nrv_patch_embed_fwd(img.data_ptr(), B, Cc, H, W, ph, pw, w.data_ptr(), kdim, bias.data_ptr(), pos.data_ptr(), D,
                                       N, cls, out.data_ptr(), D, D, sp()))
    patches = O.patchify_cp1p2(img.float().cpu(), ph, pw).reshape(B, n, kdim).double().to(dev())
    ref = patches @ w.double().t() + bias.double() + pos[cls:].double()
    o3 = out.view(B, N, D)
    assert rel(o3[:, cls:], ref) < 6e-3
    if cls:
        assert (o3[:, 0] == 7.0).all()          # the class-token rows belong to nrv_cls_token_fwd
    # weight gradient: dw += dx^T patches over the patch rows (class-token rows of dx do not take part)
    dx = torch.randn(B * N, D, generator=g).to(dev(), torch.bfloat16)
    dw = torch.ones(D, kdim, device=dev())
    _abi.check(lib.nrv_patch_embed_bwd_weight(img.data_ptr(), B, Cc, H, W, ph, pw, dx.data_ptr(), D, N, cls, dw.data_ptr(), kdim,
                                              D, sp()))
    ref_dw = torch.einsum("bnd,bnk->dk", dx.view(B, N, D)[:, cls:].double(), patches)
    assert rel(dw - 1.0, ref_dw) < 2e-3
    # unsupported layouts say so instead of computing something else
    assert lib.nrv_patch_embed_supported(Cc, H, W, ph, pw, _abi.PATCH_P1P2C, bf, bf, D) == 0
    assert lib.nrv_patch_embed_supported(3, 224, 224, 14, 14, _abi.PATCH_CP1P2, bf, bf, 1280) == 0


def test_posemb_sincos_matches_golden(golden_dir):
    import numpy as np, os
    lib = _abi.init(dev())
    pe = torch.empty(64, 512, device=dev())
    _abi.check(lib.nrv_posemb_sincos_2d(pe.data_ptr(), 8, 8, 512, 10000.0, sp()))
    gold = np.load(os.path.join(golden_dir, "posemb_sincos.npz"))["pe"]
    assert np.abs(pe.cpu().numpy() - gold).max() < 2e-6
    assert lib.nrv_posemb_sincos_2d(pe.data_ptr(), 8, 8, 510, 10000.0, sp()) != 0


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("pool", [_abi.POOL_MEAN, _abi.POOL_CLS])
def test_pool_fwd_bwd(dtype, pool):
    lib = _abi.init(dev())
    B, N, D = 5, 17, 72
    x = torch.randn(B, N, D, device=dev()).to(dtype)
    pooled = torch.empty(B, D, device=dev(), dtype=dtype)
    code = _abi._dt(x)
    _abi.check(lib.nrv_pool_fwd(x.data_ptr(), pooled.data_ptr(), B, N, D, pool, code, sp()))
    ref = x.double().mean(1) if pool == _abi.POOL_MEAN else x.double()[:, 0]
    assert rel(pooled, ref) < tol(dtype, 1e-6)
    dp = torch.randn(B, D, device=dev()).to(dtype)
    dx = torch.full((B, N, D), float("nan"), device=dev(), dtype=dtype)
    _abi.check(lib.nrv_pool_bwd(dp.data_ptr(), dx.data_ptr(), B, N, D, pool, code, sp()))
    want = torch.zeros(B, N, D, dtype=torch.double, device=dev())
    if pool == _abi.POOL_MEAN:
        want += dp.double()[:, None, :] / N
    else:
        want[:, 0] = dp.double()
    assert rel(dx, want) < tol(dtype, 1e-6)


@pytest.mark.parametrize("ls", [0.0, 0.1, 0.8])
@pytest.mark.parametrize("B,Cn", [(1, 10), (37, 100), (256, 1000)])
def test_softmax_ce(ls, B, Cn):
    import vit_pytorch_robust as v
    g = torch.Generator().manual_seed(B + Cn)
    z = (3 * torch.randn(B, Cn, generator=g)).to(dev()).requires_grad_(True)
    y = torch.randint(0, Cn, (B,), generator=g).to(dev())
    loss = v.softmax_cross_entropy(z, y, ls)
    loss.backward()
    zr = z.detach().double().requires_grad_(True)
    lr = torch.nn.functional.cross_entropy(zr, y, label_smoothing=ls)
    lr.backward()
    assert abs(loss.item() - lr.item()) < 1e-5 * max(1.0, abs(lr.item()))
    assert rel(z.grad, zr.grad) < 1e-5


# ---------------------------------------------------------------- attention
def torch_attention(qkv, B, N, H, dh, scale):
    q, k, v = qkv.view(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * scale
    p = s.softmax(-1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B, N, H * dh), torch.logsumexp(s, -1)


def _attn_case(dtype, B, N, H, dh, impl):
    lib = _abi.init(dev())
    g = torch.Generator().manual_seed(N * H + dh)
    qkv = torch.randn(B, N, 3 * H * dh, generator=g).to(dev(), dtype)
    dout = torch.randn(B, N, H * dh, generator=g).to(dev(), dtype)
    out = torch.full((B, N, H * dh), float("nan"), device=dev(), dtype=dtype)
    lse = torch.full((B, H, N), float("nan"), device=dev())
    scale = dh ** -0.5
    code = _abi._dt(qkv)
    _abi.check(lib.nrv_attn_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, dh, scale,
                                _abi.ATTN_SOFTMAX, code, impl, None, 0, sp()))
    qd = qkv.double().requires_grad_(True)
    ref, lse_ref = torch_attention(qd, B, N, H, dh, scale)
    assert rel(out, ref) < tol(dtype, 1e-5)
    assert rel(lse, lse_ref) < 1e-5
    ref.backward(dout.double())
    dqkv = torch.full_like(qkv, float("nan"))
    nb = lib.nrv_attn_bwd_workspace(B, N, H, dh)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev())
    _abi.check(lib.nrv_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
                                B, N, H, dh, scale, _abi.ATTN_SOFTMAX, code, impl, ws.data_ptr(), nb, sp()))
    torch.cuda.synchronize()
    g3 = qd.grad.view(B, N, 3, H * dh)
    d3 = dqkv.view(B, N, 3, H * dh)
    for i, nm in enumerate("qkv"):
        assert rel(d3[:, :, i], g3[:, :, i]) < tol(dtype, 2e-5, 1.5e-2), "d%s" % nm


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("B,N,H,dh", [(2, 16, 2, 32), (3, 65, 4, 64), (2, 197, 3, 64), (1, 50, 2, 80)])
def test_attention_simt_fwd_bwd(dtype, B, N, H, dh):
    _attn_case(dtype, B, N, H, dh, _abi.ATTN_IMPL_SIMT)


@pytest.mark.parametrize("B,N,H", [(1, 16, 1), (2, 64, 2), (3, 65, 4), (2, 128, 2), (2, 197, 3), (5, 208, 2), (40, 197, 12),
                                   (3, 33, 2), (7, 129, 3), (4, 192, 2), (5, 80, 1), (700, 64, 1), (101, 145, 4),
                                   (3, 224, 2), (2, 256, 3), (5, 241, 1)])   # > 208 tokens: general forward + fused backward
def test_attention_tcgen05_fwd_bwd(B, N, H):
    _attn_case(torch.bfloat16, B, N, H, 64, _abi.ATTN_IMPL_TC)


@pytest.mark.parametrize("B,N,H,dh", [(2, 257, 3, 80), (3, 197, 2, 80), (2, 300, 2, 64), (1, 384, 1, 128), (4, 50, 2, 32),
                                      (160, 257, 2, 80), (2, 129, 2, 96), (3, 16, 1, 16)])
def test_attention_tcgen05_general_forward(B, N, H, dh):
    """Forward-only tcgen05 kernel for head dims / lengths outside the training kernel (ViT-H/14: dh 80, N 257)."""
    lib = _abi.init(dev())
    g = torch.Generator().manual_seed(N * H + dh)
    qkv = torch.randn(B, N, 3 * H * dh, generator=g).to(dev(), torch.bfloat16)
    out = torch.full((B, N, H * dh), float("nan"), device=dev(), dtype=torch.bfloat16)
    lse = torch.full((B, H, N), float("nan"), device=dev())
    scale = dh ** -0.5
    if N <= 208 and dh == 64:
        pytest.skip("covered by the training kernel")
    _abi.check(lib.nrv_attn_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, dh, scale,
                                _abi.ATTN_SOFTMAX, _abi._dt(qkv), _abi.ATTN_IMPL_TC, None, 0, sp()))
    ref, lse_ref = torch_attention(qkv.double(), B, N, H, dh, scale)
    assert out.isfinite().all()
    assert rel(out, ref) < tol(torch.bfloat16, 1e-5)
    assert rel(lse, lse_ref) < 1e-5


@pytest.mark.parametrize("B,N,H,dh", [(2, 257, 3, 80), (3, 197, 2, 80), (2, 300, 2, 64), (1, 577, 2, 64), (4, 50, 2, 32),
                                      (40, 257, 16, 80), (2, 129, 2, 48), (3, 16, 1, 16), (2, 385, 1, 80), (1, 1024, 1, 64),
                                      (2, 128, 2, 80), (160, 1, 1, 80)])
def test_attention_tcgen05_general_backward(B, N, H, dh):
    """General tcgen05 backward (attention_bwd_big.cu) for head dims / lengths outside the fused training kernel
    (ViT-H/14: dh 80, N 257; 384-pixel models: N 577): all three gradients against fp64 autograd."""
    lib = _abi.init(dev())
    g = torch.Generator().manual_seed(N * H + dh)
    qkv = torch.randn(B, N, 3 * H * dh, generator=g).to(dev(), torch.bfloat16)
    dout = torch.randn(B, N, H * dh, generator=g).to(dev(), torch.bfloat16)
    out = torch.full((B, N, H * dh), float("nan"), device=dev(), dtype=torch.bfloat16)
    lse = torch.full((B, H, N), float("nan"), device=dev())
    scale = dh ** -0.5
    code = _abi._dt(qkv)
    qd = qkv.double().requires_grad_(True)
    ref, lse_ref = torch_attention(qd, B, N, H, dh, scale)
    if N <= 384:
        _abi.check(lib.nrv_attn_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, dh, scale,
                                    _abi.ATTN_SOFTMAX, code, _abi.ATTN_IMPL_AUTO, None, 0, sp()))
    else:   # beyond the forward kernels' 384 tokens: the backward only needs out and lse, whoever computed them
        out.copy_(ref.detach())
        lse.copy_(lse_ref.detach())
    ref.backward(dout.double())
    nb = lib.nrv_attn_bwd_workspace(B, N, H, dh)
    ws = torch.full((nb,), 0xFF, dtype=torch.uint8, device=dev())      # NaN patterns: the scratch must not be read before it is written
    g3 = qd.grad.view(B, N, 3, H * dh)
    runs = []
    for _ in range(2):
        dqkv = torch.full_like(qkv, float("nan"))
        _abi.check(lib.nrv_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
                                    B, N, H, dh, scale, _abi.ATTN_SOFTMAX, code, _abi.ATTN_IMPL_TC, ws.data_ptr(), nb, sp()))
        torch.cuda.synchronize()
        assert dqkv.isfinite().all()
        d3 = dqkv.view(B, N, 3, H * dh)
        for i, nm in enumerate("qkv"):
            if N == 1 and i < 2:     # a single key: P = 1, dS = 0 exactly; the kernel's dP - delta is bf16 rounding noise of O
                assert d3[:, :, i].abs().max() < 2e-2 * g3[:, :, 2].abs().max(), "d%s" % nm
                continue
            assert rel(d3[:, :, i], g3[:, :, i]) < tol(torch.bfloat16, 2e-5, 1.5e-2), "d%s" % nm
        runs.append(dqkv)
    assert torch.equal(runs[0], runs[1])       # no atomics anywhere: bit-reproducible
    if dh == 64 and N <= 256:
        return
    # a missing workspace is an error, not a fallback
    rc = lib.nrv_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
                          B, N, H, dh, scale, _abi.ATTN_SOFTMAX, code, _abi.ATTN_IMPL_TC, None, 0, sp())
    assert rc != 0


def torch_sinkhorn_attention(qkv, B, N, H, dh, scale):
    import vit_oracle as O
    q, k, v = qkv.view(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
    p = O.sinkhorn3(((q @ k.transpose(-1, -2)) * scale).softmax(-1))
    return (p @ v).permute(0, 2, 1, 3).reshape(B, N, H * dh)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("B,N,H,dh", [(2, 16, 2, 32), (3, 65, 4, 64), (2, 197, 3, 64), (150, 17, 1, 64),
                                      (2, 257, 2, 80), (1, 300, 1, 64), (170, 210, 1, 64)])   # > ~204 tokens: matrix in the L2-resident scratch
def test_sinkhorn_attention_fwd_bwd(dtype, B, N, H, dh):
    """robust=True: softmax + 3 Sinkhorn iterations (utils.py:1031-1037), forward and backward, any token count (ViT-H/14: 257)."""
    lib = _abi.init(dev())
    g = torch.Generator().manual_seed(N + H)
    qkv = torch.randn(B, N, 3 * H * dh, generator=g).to(dev(), dtype)
    dout = torch.randn(B, N, H * dh, generator=g).to(dev(), dtype)
    out = torch.full((B, N, H * dh), float("nan"), device=dev(), dtype=dtype)
    stats = torch.empty(lib.nrv_attn_stats_elems(B, N, H, _abi.ATTN_SINKHORN3), device=dev())
    scale = dh ** -0.5
    code = _abi._dt(qkv)
    nf = lib.nrv_attn_fwd_workspace(B, N, H, dh, _abi.ATTN_SINKHORN3)
    assert (nf > 0) == (N > 204)
    wf = torch.empty(max(nf, 16), dtype=torch.uint8, device=dev())
    _abi.check(lib.nrv_attn_fwd(qkv.data_ptr(), out.data_ptr(), stats.data_ptr(), B, N, H, dh, scale,
                                _abi.ATTN_SINKHORN3, code, _abi.ATTN_IMPL_AUTO, wf.data_ptr(), nf, sp()))
    qd = qkv.double().requires_grad_(True)
    ref = torch_sinkhorn_attention(qd, B, N, H, dh, scale)
    assert rel(out, ref) < tol(dtype, 2e-5)
    ref.backward(dout.double())
    dqkv = torch.full_like(qkv, float("nan"))
    nb = lib.nrv_attn_bwd_workspace(B, N, H, dh)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev())
    _abi.check(lib.nrv_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), stats.data_ptr(), dqkv.data_ptr(),
                                B, N, H, dh, scale, _abi.ATTN_SINKHORN3, code, _abi.ATTN_IMPL_AUTO, ws.data_ptr(), nb,
                                sp()))
    torch.cuda.synchronize()
    assert rel(dqkv, qd.grad) < tol(dtype, 1e-4, 2e-2)


@pytest.mark.parametrize("B,N,H", [(1, 1, 1), (2, 16, 2), (3, 65, 4), (4, 128, 2), (3, 129, 1), (2, 197, 3), (2, 208, 2), (40, 197, 12),
                                   (300, 64, 2)])
def test_sinkhorn_attention_tcgen05_forward_backward(B, N, H):
    """robust=True on the tensor cores (attention_sinkhorn_tc.cu: scaling-vector form, the probabilities as bf16 in shared
    memory): forward against fp64 torch, its statistics against the CUDA-core kernel's (same [B,H,8,N] layout), and both
    backward kernels -- tcgen05 (closed-form gradient through the scaling vectors) and CUDA cores (step-by-step) -- fed with
    the tcgen05 forward's statistics, against autograd."""
    lib = _abi.init(dev())
    dh = 64
    g = torch.Generator().manual_seed(7 * N + H)
    qkv = torch.randn(B, N, 3 * H * dh, generator=g).to(dev(), torch.bfloat16)
    dout = torch.randn(B, N, H * dh, generator=g).to(dev(), torch.bfloat16)
    scale = dh ** -0.5
    outs, stats = {}, {}
    nf = lib.nrv_attn_fwd_workspace(B, N, H, dh, _abi.ATTN_SINKHORN3)     # the CUDA-core kernel above ~204 tokens
    wf = torch.empty(max(nf, 16), dtype=torch.uint8, device=dev())
    for impl in (_abi.ATTN_IMPL_TC, _abi.ATTN_IMPL_SIMT):
        outs[impl] = torch.full((B, N, H * dh), float("nan"), device=dev(), dtype=torch.bfloat16)
        stats[impl] = torch.full((B, H, 8, N), float("nan"), device=dev())
        _abi.check(lib.nrv_attn_fwd(qkv.data_ptr(), outs[impl].data_ptr(), stats[impl].data_ptr(), B, N, H, dh, scale,
                                    _abi.ATTN_SINKHORN3, _abi.NRV_BF16, impl, wf.data_ptr(), nf, sp()))
    qd = qkv.double().requires_grad_(True)
    ref = torch_sinkhorn_attention(qd, B, N, H, dh, scale)
    assert rel(outs[_abi.ATTN_IMPL_TC], ref) < 8e-3
    assert rel(outs[_abi.ATTN_IMPL_SIMT], ref) < 6e-3
    assert rel(stats[_abi.ATTN_IMPL_TC], stats[_abi.ATTN_IMPL_SIMT]) < 5e-3
    ref.backward(dout.double())
    nb = lib.nrv_attn_bwd_workspace(B, N, H, dh)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev())
    g3 = qd.grad.view(B, N, 3, H * dh)
    for impl in (_abi.ATTN_IMPL_TC, _abi.ATTN_IMPL_SIMT):
        dqkv = torch.full_like(qkv, float("nan"))
        _abi.check(lib.nrv_attn_bwd(qkv.data_ptr(), outs[_abi.ATTN_IMPL_TC].data_ptr(), dout.data_ptr(),
                                    stats[_abi.ATTN_IMPL_TC].data_ptr(), dqkv.data_ptr(), B, N, H, dh, scale, _abi.ATTN_SINKHORN3,
                                    _abi.NRV_BF16, impl, ws.data_ptr(), nb, sp()))
        torch.cuda.synchronize()
        d3 = dqkv.view(B, N, 3, H * dh)
        for i, nm in enumerate("qkv"):
            assert rel(d3[:, :, i], g3[:, :, i]) < 2e-2, (impl, "d" + nm)


@pytest.mark.parametrize("mode", [_abi.ATTN_SOFTMAX, _abi.ATTN_SINKHORN3])
@pytest.mark.parametrize("B,N,H,dh", [(2, 65, 3, 64), (2, 197, 2, 64), (1, 257, 2, 80)])
def test_attention_probabilities_for_introspection(mode, B, N, H, dh):
    """nrv_attn_probs: the [B,H,N,N] matrix the reference's `attend` module returns (recorder.py:28-31), both attention modes."""
    lib = _abi.init(dev())
    g = torch.Generator().manual_seed(N + H + mode)
    qkv = torch.randn(B, N, 3 * H * dh, generator=g).to(dev(), torch.bfloat16)
    probs = torch.full((B, H, N, N), float("nan"), device=dev())
    stats = torch.empty(lib.nrv_attn_stats_elems(B, N, H, _abi.ATTN_SINKHORN3), device=dev())
    nf = lib.nrv_attn_fwd_workspace(B, N, H, dh, _abi.ATTN_SINKHORN3)
    wf = torch.empty(max(nf, 16), dtype=torch.uint8, device=dev())
    scale = dh ** -0.5
    _abi.check(lib.nrv_attn_probs(qkv.data_ptr(), probs.data_ptr(), stats.data_ptr(), B, N, H, dh, scale, mode, _abi.NRV_BF16,
                                  wf.data_ptr(), nf, sp()))
    q, k, v = qkv.double().view(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
    p = ((q @ k.transpose(-1, -2)) * scale).softmax(-1)
    if mode == _abi.ATTN_SINKHORN3:
        for _ in range(3):
            p = p / p.sum(-1, keepdim=True)
            p = p / p.sum(-2, keepdim=True)
        p = p / p.sum(-1, keepdim=True)
    assert rel(probs, p) < 2e-5


def test_sinkhorn_needs_its_scratch_for_long_sequences_and_rejects_oversized_heads():
    lib = _abi.init(dev())
    t = torch.zeros(1, 257, 3 * 80, device=dev(), dtype=torch.bfloat16)
    o = torch.zeros(1, 257, 80, device=dev(), dtype=torch.bfloat16)
    st = torch.zeros(8 * 257, device=dev())
    rc = lib.nrv_attn_fwd(t.data_ptr(), o.data_ptr(), st.data_ptr(), 1, 257, 1, 80, 0.1, _abi.ATTN_SINKHORN3, 0, 0, None, 0, sp())
    assert rc == -1 and b"scratch" in lib.nrv_last_error()   # NRV_EINVAL: no silent fallback
    # the K / V tile of one head must still fit in shared memory: 4000 tokens x 64 do not
    assert lib.nrv_attn_fwd_workspace(1, 4000, 1, 64, _abi.ATTN_SINKHORN3) > 0
    big = torch.zeros(1, 4000, 3 * 64, device=dev(), dtype=torch.bfloat16)
    ob = torch.zeros(1, 4000, 64, device=dev(), dtype=torch.bfloat16)
    sb = torch.zeros(8 * 4000, device=dev())
    rc = lib.nrv_attn_fwd(big.data_ptr(), ob.data_ptr(), sb.data_ptr(), 1, 4000, 1, 64, 0.1, _abi.ATTN_SINKHORN3, 0, 0, None, 0, sp())
    assert rc == -5  # NRV_ENOTIMPL


# ---------------------------------------------------------------- optimiser pieces
def test_adamw_matches_torch():
    lib = _abi.init(dev())
    n = 100003
    g0 = torch.Generator().manual_seed(1)
    p = torch.randn(n, generator=g0).to(dev())
    ref = torch.nn.Parameter(p.clone().double())
    opt = torch.optim.AdamW([ref], lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    shadow = torch.empty(n, device=dev(), dtype=torch.bfloat16)
    for step in range(1, 6):
        g = torch.randn(n, generator=g0).to(dev())
        ref.grad = g.double().clone()
        opt.step()
        _abi.check(lib.nrv_adamw(p.data_ptr(), m.data_ptr(), v.data_ptr(), g.data_ptr(), shadow.data_ptr(), n,
                                 2e-4, 0.9, 0.999, 1e-8, 0.01, step, 1.0, None, sp()))
    assert rel(p, ref.detach()) < 1e-6
    assert torch.equal(shadow, p.to(torch.bfloat16))


def test_sumsq_and_clip():
    lib = _abi.init(dev())
    g = torch.randn(70001, device=dev())
    acc = torch.zeros(2, device=dev())
    _abi.check(lib.nrv_sumsq(g.data_ptr(), g.numel(), acc.data_ptr(), sp()))
    assert abs(acc[0].item() - (g.double() ** 2).sum().item()) < 1e-3 * g.numel() ** 0.5
    _abi.check(lib.nrv_clip_coef(acc.data_ptr(), 5.0, 1.0, acc.data_ptr() + 4, sp()))
    want = min(1.0, 5.0 / (g.double().norm().item() + 1e-6))
    assert abs(acc[1].item() - want) < 1e-6


@pytest.mark.parametrize("dtype", DT)
def test_add_gaussian_noise(dtype):
    """examples/nowak.py:153  x + std * randn_like(x): first two moments, independence from x, determinism by seed."""
    import vit_pytorch_robust as V
    n = 1 << 22
    x = torch.linspace(-1, 1, n, device=dev()).to(dtype)
    y = V.add_gaussian_noise(x, 0.1, seed=7)
    d = (y.double() - x.double())
    tol = 2e-3 if dtype == torch.bfloat16 else 3e-4          # bf16 rounding of x + noise adds its own jitter
    assert abs(d.mean().item()) < tol
    assert abs(d.std().item() - 0.1) < tol
    z = d / 0.1
    assert abs((z ** 3).mean().item()) < 2e-2 and abs((z ** 4).mean().item() - 3.0) < 5e-2   # Gaussian shape
    assert abs((z[:-1] * z[1:]).mean().item()) < 3e-3                                          # neighbours uncorrelated
    assert abs((z * x.double()).mean().item()) < 3e-3                                          # independent of x
    assert torch.equal(y, V.add_gaussian_noise(x, 0.1, seed=7))
    y2 = V.add_gaussian_noise(x, 0.1, seed=8)
    assert abs(((y2.double() - x.double()) * d).mean().item()) < 1e-4
    torch.manual_seed(5)
    a = V.add_gaussian_noise(x, 0.1)
    torch.manual_seed(5)
    assert torch.equal(a, V.add_gaussian_noise(x, 0.1))


@pytest.mark.parametrize("M,N,K", [(200, 136, 72), (1000, 776, 328), (33, 8, 64), (4100, 3072, 128)])
def test_gemm_mul_epilogue_fused_colsum_ragged_shapes(M, N, K):
    """EPI_MUL + colsum (the fc1 bias gradient) on shapes whose last tiles are partial in M and N: rows beyond M and
    columns beyond N must contribute nothing (TMA zero-fills the operand and factor tiles; the reds are column-guarded)."""
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(dev(), torch.bfloat16)
    B = (torch.randn(N, K, generator=g) / 8).to(dev(), torch.bfloat16)
    aux = torch.randn(M, N, generator=g).to(dev(), torch.bfloat16)
    out = torch.empty(M, N, device=dev(), dtype=torch.bfloat16)
    cs = torch.zeros(N, device=dev(), dtype=torch.float32)
    _abi.gemm(A, B, out, b_layout=_abi.NRV_K_MAJOR, epi=_abi.EPI_MUL, aux=aux, colsum=cs)
    ref = (A.double() @ B.double().t()) * aux.double()
    assert rel(out, ref) < tol(torch.bfloat16)
    assert rel(cs, out.double().sum(0)) < 1e-5
