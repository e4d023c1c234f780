"""Oracle parity AT THE SIZES OF THE BASELINE.json CONFIGS (VERDICT r1 "what's weak" #1): the full-depth models, every
layer, real token counts and widths, small batches so the fp64 CPU oracle finishes in seconds.

  configs[0]  README ViT / SimpleViT 256^2, patch 32, dim 1024, depth 6, heads 16, mlp 2048, 1000 classes (README.md:67-86,122-139), B=8
  configs[1]  SimpleViT CIFAR-100 shape 32^2, patch 4, dim 512, depth 6, heads 8, mlp 2048, B=8
  configs[2]  vit_b_16()  (12 layers, 197 tokens), B=4           <- the headline model of bench.py
  configs[3]  vit_l_16()  (24 layers), B=2
  configs[4]  vit_h_14()  (32 layers, 257 tokens, dh=80), forward only, B=2

Checked: logits and EVERY parameter gradient.  Tolerances are BASELINE.json's: rel-L2 <= 1e-3 in the fp32 check mode,
cosine >= 0.999 in bf16 (twelve to thirty-two layers of bf16 residual stream, the sigmoid-form GELU / stored GELU',
fp32-atomic split-K dW and the fused bias-gradient epilogues all measured together)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import vit_oracle as O  # noqa: E402
import vit_pytorch_robust as V  # noqa: E402
from helpers import compare_grads, model_loss_and_grads, randomize_  # noqa: E402

DEV = "cuda:0"
CHECK_REL = 1e-3      # north_star: rel <= 1e-3 in the FP32-accumulate check mode
BF16_COS = 0.999      # north_star: cosine >= 0.999 for BF16


def _vit_ref(sd, img, labels, patch, heads, forward_only=False):
    fwd = lambda s, x: O.vision_transformer_forward(s, x, patch_size=patch, num_heads=heads)  # noqa: E731
    if forward_only:
        with torch.no_grad():
            return fwd({k: v.double() for k, v in sd.items()}, img.double()), None, None
    return O.loss_and_grads(fwd, sd, img.double(), labels, 0.1)


def _check(m, dtype, img, labels, ref_logits, ref_grads):
    m._nrv.compute_dtype = dtype
    lg, ls, gr = model_loss_and_grads(m, img.to(DEV), labels.to(DEV), 0.1)
    assert set(gr) == set(ref_grads)
    if dtype == torch.float32:
        assert O.rel_l2(lg, ref_logits) < CHECK_REL
        worst, key = compare_grads(gr, ref_grads, O.rel_l2)
        assert worst < CHECK_REL, (key, worst)
    else:
        assert O.cosine(lg, ref_logits) > BF16_COS
        worst, key = compare_grads(gr, ref_grads, O.cosine)
        assert worst > BF16_COS, (key, worst)
    return worst, key


def _prep(m, seed, B, size, ncls):
    randomize_(m, seed)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(seed + 1)
    img = torch.randn(B, 3, size, size, generator=g)
    labels = torch.randint(0, ncls, (B,), generator=g)
    return sd, img, labels


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["check_fp32", "bf16"])
def test_vit_b16_full_model_matches_oracle(dtype):
    """BASELINE configs[2]: vit_b_16(), all 12 layers, 224^2 -> 197 tokens, B=4."""
    torch.manual_seed(0)
    m = V.vit_b_16()
    sd, img, labels = _prep(m, 101, 4, 224, 1000)
    ref_logits, _, ref_grads = _vit_ref(sd, img, labels, 16, 12)
    _check(m.to(DEV), dtype, img, labels, ref_logits, ref_grads)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["check_fp32", "bf16"])
def test_vit_l16_full_model_matches_oracle(dtype):
    """BASELINE configs[3]: vit_l_16(), 24 layers, B=2."""
    torch.manual_seed(0)
    m = V.vit_l_16()
    sd, img, labels = _prep(m, 103, 2, 224, 1000)
    ref_logits, _, ref_grads = _vit_ref(sd, img, labels, 16, 16)
    _check(m.to(DEV), dtype, img, labels, ref_logits, ref_grads)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["check_fp32", "bf16"])
def test_vit_h14_full_model_forward_matches_oracle(dtype):
    """BASELINE configs[4]: vit_h_14() inference, 32 layers, 257 tokens, dh = 80, patch_dim 588, B=2 (eval + no_grad)."""
    torch.manual_seed(0)
    m = V.vit_h_14()
    sd, img, labels = _prep(m, 105, 2, 224, 1000)
    ref_logits, _, _ = _vit_ref(sd, img, labels, 14, 16, forward_only=True)
    m = m.to(DEV).eval()
    m._nrv.compute_dtype = dtype
    with torch.no_grad():
        lg = m(img.to(DEV)).float().cpu()
    if dtype == torch.float32:
        assert O.rel_l2(lg, ref_logits) < CHECK_REL
    else:
        assert O.cosine(lg, ref_logits) > BF16_COS


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["check_fp32", "bf16"])
@pytest.mark.parametrize("kind", ["SimpleViT", "ViT"])
def test_readme_config_matches_oracle(kind, dtype):
    """BASELINE configs[0]: README.md:67-86 (ViT) / :122-139 (SimpleViT): 256^2, patch 32, dim 1024, depth 6, heads 16,
    mlp 2048, 1000 classes, batch 8."""
    kw = dict(image_size=256, patch_size=32, num_classes=1000, dim=1024, depth=6, heads=16, mlp_dim=2048)
    torch.manual_seed(0)
    m = V.SimpleViT(**kw) if kind == "SimpleViT" else V.ViT(**kw)
    sd, img, labels = _prep(m, 107, 8, 256, 1000)
    if kind == "SimpleViT":
        fwd = lambda s, x: O.simple_vit_forward(s, x, patch_size=32, heads=16, dim_head=64)  # noqa: E731
    else:
        fwd = lambda s, x: O.readme_vit_forward(s, x, patch_size=32, heads=16, dim_head=64)  # noqa: E731
    ref_logits, _, ref_grads = O.loss_and_grads(fwd, sd, img.double(), labels, 0.1)
    _check(m.to(DEV), dtype, img, labels, ref_logits, ref_grads)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["check_fp32", "bf16"])
@pytest.mark.parametrize("robust", [False, True], ids=["softmax", "robust"])
def test_cifar_simplevit_config_matches_oracle(robust, dtype):
    """BASELINE configs[1]: SimpleViT CIFAR-100 shape (32^2, patch 4, dim 512, depth 6, heads 8, mlp 2048), B=8,
    with softmax and with robust=True (SinkhornAttention, simple_vit.py:56-57)."""
    kw = dict(image_size=32, patch_size=4, num_classes=100, dim=512, depth=6, heads=8, mlp_dim=2048)
    torch.manual_seed(0)
    m = V.SimpleViT(**kw, robust=robust)
    sd, img, labels = _prep(m, 109, 8, 32, 100)
    fwd = lambda s, x: O.simple_vit_forward(s, x, patch_size=4, heads=8, dim_head=64, robust=robust)  # noqa: E731
    ref_logits, _, ref_grads = O.loss_and_grads(fwd, sd, img.double(), labels, 0.1)
    _check(m.to(DEV), dtype, img, labels, ref_logits, ref_grads)
