"""Fused softmax cross-entropy with label smoothing (nrv_softmax_ce): loss and dlogits in one pass
over the logits — replaces F.cross_entropy(preds, y, label_smoothing=...) (examples/baseline.py:70);
add_gaussian_noise: the examples' noisy-input objective x + std * randn_like(x) (examples/nowak.py:153) in one pass."""
import torch

from . import _abi


class _SoftmaxCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, label_smoothing):
        if not logits.is_cuda:
            raise _abi.NrvError("softmax_cross_entropy needs CUDA logits (no CPU fallback)")
        lib = _abi.init(logits.device)
        lg = logits.float().contiguous() if (logits.dtype != torch.float32 or logits.stride(1) != 1) else logits
        B, Cn = lg.shape
        loss = torch.empty((), dtype=torch.float32, device=lg.device)
        dl = torch.empty(B, Cn, dtype=torch.float32, device=lg.device)
        lab = labels.to(torch.int64).contiguous()
        _abi.check(lib.nrv_softmax_ce(lg.data_ptr(), lg.stride(0), lab.data_ptr(), float(label_smoothing),
                                      loss.data_ptr(), dl.data_ptr(), _abi.NRV_F32, Cn, 1.0, B, Cn,
                                      _abi.stream_ptr()), "nrv_softmax_ce")
        ctx.save_for_backward(dl)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g, None, None


def softmax_cross_entropy(logits, labels, label_smoothing=0.0):
    """Mean cross-entropy over the batch; logits [B, C] fp32, labels [B] int64."""
    return _SoftmaxCE.apply(logits, labels, label_smoothing)


def add_gaussian_noise(x, std=0.1, seed=None):
    """x + std * randn_like(x) in one kernel (bf16 or fp32 CUDA tensor, numel % 8 == 0).  The seed defaults to a draw from
    torch's default CPU generator, so torch.manual_seed reproduces it; no gradient flows to the noise."""
    if not x.is_cuda:
        raise _abi.NrvError("add_gaussian_noise needs a CUDA tensor (no CPU fallback)")
    if x.dtype not in (torch.float32, torch.bfloat16) or x.numel() % 8 != 0:
        raise ValueError("add_gaussian_noise: bf16 / fp32 tensor with numel % 8 == 0 expected")
    lib = _abi.init(x.device)
    xc = x.contiguous()
    out = torch.empty_like(xc)
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    import ctypes as C
    _abi.check(lib.nrv_add_gaussian_noise(xc.data_ptr(), out.data_ptr(), xc.numel(), _abi._dt(xc), float(std),
                                          C.c_ulonglong(seed), _abi.stream_ptr()), "nrv_add_gaussian_noise")
    return out

