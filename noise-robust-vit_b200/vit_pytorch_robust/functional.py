"""Fused softmax cross-entropy with label smoothing (nrv_softmax_ce): loss and dlogits in one pass
over the logits — replaces F.cross_entropy(preds, y, label_smoothing=...) (examples/baseline.py:70)."""
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
