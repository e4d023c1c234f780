"""GPU diagnostic for nrv_gemm: error statistics for every operand-layout / epilogue variant and
CUDA-event timings at the ViT-B/16 hot shapes.  Run on the B200 box:
    python tools/gpu_check_gemm.py [--quick]
Prints one line per case; exits non-zero if any case is out of tolerance."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
from vit_pytorch_robust import _abi  # noqa: E402

dev = torch.device("cuda:0")
fails = []


def rel(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def report(name, got, ref, tol):
    r = rel(got, ref)
    mx = (got.double() - ref.double()).abs().max().item()
    ok = r <= tol and got.isfinite().all().item()
    print("%-58s rel %.3e max %.3e %s" % (name, r, mx, "ok" if ok else "FAIL"), flush=True)
    if not ok:
        fails.append(name)
        # locate the damage: per 128x64 block error map summary
        d = (got.double() - ref.double()).abs()
        bad = (d > 10 * tol * ref.double().abs().max()).nonzero()
        if bad.numel():
            print("    first bad idx", bad[:4].tolist(), "count", bad.shape[0], "of", d.numel(), flush=True)


def run_case(M, N, K, a_mn, b_mn, dtype, **kw):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K + a_mn * 2 + b_mn)
    A = torch.randn(M, K, generator=g).to(dev, dtype)
    B = torch.randn(N, K, generator=g).to(dev, dtype)
    a_arg = A.t().contiguous() if a_mn else A
    b_arg = B.t().contiguous() if b_mn else B
    out = torch.full((M, N), float("nan"), device=dev, dtype=dtype)
    _abi.gemm(a_arg, b_arg, out, a_layout=int(a_mn), b_layout=int(b_mn), M=M, N=N, K=K, **kw)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    return out, ref


def main():
    quick = "--quick" in sys.argv
    torch.manual_seed(0)
    _abi.init(dev)
    print("device", torch.cuda.get_device_name(0), "sms", _abi.load().nrv_num_sms(), flush=True)

    # ---- 1. plain K-major bf16, small → large, ragged edges
    for (M, N, K) in [(128, 128, 64), (128, 256, 64), (128, 256, 128), (256, 512, 256), (200, 136, 72),
                      (1000, 1000, 768), (4096, 2304, 768), (333, 3072, 768), (512, 768, 3072)]:
        try:
            out, ref = run_case(M, N, K, 0, 0, torch.bfloat16)
            report("bf16 KK  %dx%dx%d" % (M, N, K), out, ref, 6e-3)
        except Exception as e:  # noqa: BLE001
            print("bf16 KK %dx%dx%d EXC %s" % (M, N, K, e), flush=True)
            fails.append("exc")
            break
    # force BN=128 tile
    out, ref = run_case(512, 768, 768, 0, 0, torch.bfloat16, force_bn128=1)
    report("bf16 KK  512x768x768 bn128", out, ref, 6e-3)

    # ---- 2. MN-major operands (dX: B MN ; dW: A MN, B MN)
    for (a_mn, b_mn) in [(0, 1), (1, 0), (1, 1)]:
        for (M, N, K) in [(128, 256, 64), (256, 512, 256), (768, 768, 4096), (200, 136, 72)]:
            try:
                out, ref = run_case(M, N, K, a_mn, b_mn, torch.bfloat16)
                report("bf16 a_mn=%d b_mn=%d %dx%dx%d" % (a_mn, b_mn, M, N, K), out, ref, 6e-3)
            except Exception as e:  # noqa: BLE001
                print("bf16 mn case EXC %s" % e, flush=True)
                fails.append("exc")

    # ---- 3. tf32 check mode
    for (a_mn, b_mn) in [(0, 0), (0, 1), (1, 1)]:
        for (M, N, K) in [(128, 256, 32), (256, 512, 256), (200, 136, 72)]:
            try:
                out, ref = run_case(M, N, K, a_mn, b_mn, torch.float32)
                report("tf32 a_mn=%d b_mn=%d %dx%dx%d" % (a_mn, b_mn, M, N, K), out, ref, 1e-3)
            except Exception as e:  # noqa: BLE001
                print("tf32 case EXC %s" % e, flush=True)
                fails.append("exc")

    # ---- 4. epilogues
    M, N, K = 384, 512, 256
    g = torch.Generator().manual_seed(5)
    A = torch.randn(M, K, generator=g).to(dev, torch.bfloat16)
    B = (torch.randn(N, K, generator=g) / 16).to(dev, torch.bfloat16)
    bias = torch.randn(N, generator=g).to(dev)
    res = torch.randn(M, N, generator=g).to(dev, torch.bfloat16)
    acc = A.double() @ B.double().t()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    _abi.gemm(A, B, out, bias=bias, residual=res, alpha=0.5)
    report("epi bias+residual alpha", out, 0.5 * acc + bias.double() + res.double(), 6e-3)
    out2 = torch.empty_like(out)
    _abi.gemm(A, B, out, bias=bias, epi=_abi.EPI_GELU, out2=out2)
    u = acc + bias.double()
    report("epi gelu (act)", out, torch.nn.functional.gelu(u), 6e-3)
    report("epi gelu (pre-act)", out2, u, 6e-3)
    aux = torch.randn(M, N, generator=g).to(dev, torch.bfloat16)
    _abi.gemm(A, B, out, epi=_abi.EPI_DGELU, aux=aux)
    x = aux.double().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    report("epi dgelu", out, acc * x.grad, 6e-3)
    g2 = torch.empty_like(out)
    _abi.gemm(A, B, out, bias=bias, epi=_abi.EPI_GELU_GRAD, out2=g2)
    xu = u.clone().requires_grad_(True)
    torch.nn.functional.gelu(xu).sum().backward()
    report("epi gelu_grad (act)", out, torch.nn.functional.gelu(u), 6e-3)
    report("epi gelu_grad (grad)", g2, xu.grad, 6e-3)
    _abi.gemm(A, B, out, epi=_abi.EPI_MUL, aux=aux)
    report("epi mul", out, acc * aux.double(), 6e-3)
    outf = torch.zeros(M, N, device=dev, dtype=torch.float32)
    _abi.gemm(A, B, outf, epi=_abi.EPI_ATOMIC_F32, splits=4)
    _abi.gemm(A, B, outf, epi=_abi.EPI_ATOMIC_F32, splits=0)
    report("epi atomic f32 splitK x2", outf, 2 * acc, 1e-4)
    outf32 = torch.empty(M, N, device=dev, dtype=torch.float32)
    _abi.gemm(A, B, outf32, bias=bias)
    report("bf16 in, f32 out + bias", outf32, acc + bias.double(), 1e-5)
    # token-row remap + positional table (patch embed)
    Bsz, n_in, n_out = 6, 64, 65
    pos = torch.randn(n_out, N, generator=g).to(dev)
    outp = torch.zeros(Bsz * n_out, N, device=dev, dtype=torch.bfloat16)
    _abi.gemm(A, B, outp, bias=bias, pos=pos, pos_rows_in=n_in, pos_rows_out=n_out, pos_row_off=1)
    refp = torch.zeros(Bsz, n_out, N, dtype=torch.double, device=dev)
    refp[:, 1:] = (acc + bias.double()).view(Bsz, n_in, N) + pos.double()[1:]
    report("epi patch remap + pos", outp.view(Bsz, n_out, N), refp, 6e-3)

    # ---- 4b. epilogue cost at the FC1 / FC2-dX shape
    if "--epi" in sys.argv:
        M, N, K = 50432, 3072, 768
        A = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        B = torch.randn(N, K, device=dev, dtype=torch.bfloat16) / 16
        bias = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        out2 = torch.empty_like(out)
        aux = torch.randn(M, N, device=dev, dtype=torch.bfloat16)
        A2 = torch.randn(M, 768, device=dev, dtype=torch.bfloat16)
        B2 = torch.randn(768, 768, device=dev, dtype=torch.bfloat16) / 16
        o2 = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
        r2 = torch.randn(M, 768, device=dev, dtype=torch.bfloat16)
        b2 = torch.randn(768, device=dev)
        cases = [("store", lambda: _abi.gemm(A, B, out)),
                 ("store+bias", lambda: _abi.gemm(A, B, out, bias=bias)),
                 ("gelu (h only)", lambda: _abi.gemm(A, B, out, bias=bias, epi=_abi.EPI_GELU)),
                 ("gelu + u", lambda: _abi.gemm(A, B, out, bias=bias, epi=_abi.EPI_GELU, out2=out2)),
                 ("gelu_grad", lambda: _abi.gemm(A, B, out, bias=bias, epi=_abi.EPI_GELU_GRAD, out2=out2)),
                 ("dgelu", lambda: _abi.gemm(A, B, out, epi=_abi.EPI_DGELU, aux=aux)),
                 ("mul", lambda: _abi.gemm(A, B, out, epi=_abi.EPI_MUL, aux=aux)),
                 ("N768 store", lambda: _abi.gemm(A2, B2, o2)),
                 ("N768 bias+residual", lambda: _abi.gemm(A2, B2, o2, bias=b2, residual=r2))]
        for name, fn in cases:
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            nn = 768 if name.startswith("N768") else N
            print("epi-time %-20s %.3f ms  %.0f TF" % (name, ms, 2.0 * M * nn * K / ms / 1e9), flush=True)

    # ---- 5. timings at the ViT-B/16 hot shapes (B=256 -> T=50432)
    if not quick:
        shapes = [("qkv", 50432, 2304, 768, 0, 0, {}), ("attn_out", 50432, 768, 768, 0, 0, {}),
                  ("fc1", 50432, 3072, 768, 0, 0, {}), ("fc2", 50432, 768, 3072, 0, 0, {}),
                  ("dX fc1", 50432, 768, 3072, 0, 1, {}), ("dX fc2", 50432, 3072, 768, 0, 1, {}),
                  ("dW fc1", 3072, 768, 50432, 1, 1, {"atomic": 1}), ("dW qkv", 2304, 768, 50432, 1, 1, {"atomic": 1}),
                  ("dW out", 768, 768, 50432, 1, 1, {"atomic": 1})]
        for name, M, N, K, a_mn, b_mn, opt in shapes:
            A = torch.randn((K, M) if a_mn else (M, K), device=dev, dtype=torch.bfloat16)
            B = torch.randn((K, N) if b_mn else (N, K), device=dev, dtype=torch.bfloat16)
            atomic = opt.get("atomic", 0)
            out = torch.zeros(M, N, device=dev, dtype=torch.float32 if atomic else torch.bfloat16)
            kw = dict(a_layout=a_mn, b_layout=b_mn, M=M, N=N, K=K)
            if atomic:
                kw["epi"] = _abi.EPI_ATOMIC_F32
            for _ in range(3):
                _abi.gemm(A, B, out, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 10
            e0.record()
            for _ in range(iters):
                _abi.gemm(A, B, out, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            # torch reference timing (cuBLAS) for context
            At = A.t() if a_mn else A
            Bt = B if b_mn else B.t()
            for _ in range(3):
                torch.matmul(At, Bt)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                torch.matmul(At, Bt)
            e1.record()
            torch.cuda.synchronize()
            ms_ref = e0.elapsed_time(e1) / iters
            tf = 2.0 * M * N * K / ms / 1e9
            kw1 = dict(kw, force_single_cta=1)
            for _ in range(3):
                _abi.gemm(A, B, out, **kw1)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                _abi.gemm(A, B, out, **kw1)
            e1.record()
            torch.cuda.synchronize()
            ms1 = e0.elapsed_time(e1) / iters
            print("time %-9s %6dx%5dx%6d  pair %.3f ms %.0f TF | single-CTA %.3f ms %.0f TF | cuBLAS %.3f ms %.0f TF" %
                  (name, M, N, K, ms, tf, ms1, 2.0 * M * N * K / ms1 / 1e9, ms_ref, 2.0 * M * N * K / ms_ref / 1e9), flush=True)

    print("FAILS:", fails)
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
