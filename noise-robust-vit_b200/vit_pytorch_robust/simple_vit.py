"""SimpleViT — drop-in for the reference's vit_pytorch_robust/simple_vit.py.

Same constructor, same forward(img) -> logits, same module tree and state_dict keys
(reference simple_vit.py:100-149), but forward() is ONE fused call chain into libnrvit
(nrv_vit_forward + head GEMM) instead of ~14 ATen kernels per layer.  The sub-modules below only
own parameters under the reference's names; they are not executed one by one.
"""
import torch
from torch import nn

from . import engine as _engine


def pair(t):
    """simple_vit.py:11-12"""
    return t if isinstance(t, tuple) else (t, t)


class Rearrange(nn.Module):
    """Parameter-free stand-in for einops' Rearrange('b c (h p1) (w p2) -> b h w (p1 p2 c)')
    (simple_vit.py:127-129): keeps index 0 of to_patch_embedding so that the Linear stays at key
    `to_patch_embedding.1`.  The rearrangement itself is the im2col prologue of the patch GEMM."""

    def __init__(self, pattern, **axes):
        super().__init__()
        self.pattern, self.axes = pattern, axes

    def extra_repr(self):
        return "%r, %s" % (self.pattern, ", ".join("%s=%d" % kv for kv in self.axes.items()))

    def forward(self, x):
        raise NotImplementedError("Rearrange is fused into the patch-embedding kernel; call the model, not the sub-module")


class _FusedOnly(nn.Module):
    def forward(self, *a, **k):
        raise NotImplementedError(
            "%s only holds parameters: the encoder runs as one fused libnrvit call from the model's forward()" %
            type(self).__name__)


class FeedForward(_FusedOnly):
    """simple_vit.py:34-45 — LayerNorm, Linear, GELU, Linear."""

    def __init__(self, dim, hidden_dim):
        super().__init__()
        self.net = nn.Sequential(
            nn.LayerNorm(dim),
            nn.Linear(dim, hidden_dim),
            nn.GELU(),
            nn.Linear(hidden_dim, dim),
        )


class Attention(_FusedOnly):
    """simple_vit.py:48-76 — LayerNorm, to_qkv (no bias), softmax | Sinkhorn attention, to_out (no bias)."""

    def __init__(self, dim, heads=8, dim_head=64, robust=False):
        super().__init__()
        inner_dim = dim_head * heads
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.robust = robust
        self.norm = nn.LayerNorm(dim)
        self.attend = nn.Identity()  # softmax / SinkhornAttention run inside the attention kernel
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = nn.Linear(inner_dim, dim, bias=False)


class Transformer(_FusedOnly):
    """simple_vit.py:79-97"""

    def __init__(self, dim, depth, heads, dim_head, mlp_dim, robust):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                Attention(dim, heads=heads, dim_head=dim_head, robust=robust),
                FeedForward(dim, mlp_dim),
            ]))


class SimpleViT(nn.Module):
    """simple_vit.py:100-149"""

    def __init__(self, *, image_size, patch_size, num_classes, dim, depth, heads, mlp_dim,
                 channels=3, dim_head=64, robust=False):
        super().__init__()
        image_height, image_width = pair(image_size)
        patch_height, patch_width = pair(patch_size)

        assert (
            image_height % patch_height == 0 and image_width % patch_width == 0
        ), "Image dimensions must be divisible by the patch size."

        patch_dim = channels * patch_height * patch_width

        self.to_patch_embedding = nn.Sequential(
            Rearrange("b c (h p1) (w p2) -> b h w (p1 p2 c)", p1=patch_height, p2=patch_width),
            nn.Linear(patch_dim, dim),
        )
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim, robust)
        self.to_latent = nn.Identity()
        self.linear_head = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, num_classes))

        self._nrv = _engine.Engine(
            dict(image_size=(image_height, image_width), patch_size=(patch_height, patch_width),
                 channels=channels, dim=dim, depth=depth, heads=heads, dim_head=dim_head, mlp_dim=mlp_dim,
                 cls_token=False, pool="mean", patch_order="p1p2c", qkv_bias=False, ln_eps=1e-5,
                 robust=robust),
            self._nrv_param_map)

    def _nrv_param_map(self):
        pm = {
            "w_patch": self.to_patch_embedding[1].weight, "b_patch": self.to_patch_embedding[1].bias,
            "lnf_g": self.linear_head[0].weight, "lnf_b": self.linear_head[0].bias,
            "head_w": self.linear_head[1].weight, "head_b": self.linear_head[1].bias,
        }
        for i, (attn, ff) in enumerate(self.transformer.layers):
            pre = "l%d." % i
            pm[pre + "ln1_g"], pm[pre + "ln1_b"] = attn.norm.weight, attn.norm.bias
            pm[pre + "w_qkv"], pm[pre + "w_out"] = attn.to_qkv.weight, attn.to_out.weight
            pm[pre + "ln2_g"], pm[pre + "ln2_b"] = ff.net[0].weight, ff.net[0].bias
            pm[pre + "w_fc1"], pm[pre + "b_fc1"] = ff.net[1].weight, ff.net[1].bias
            pm[pre + "w_fc2"], pm[pre + "b_fc2"] = ff.net[3].weight, ff.net[3].bias
        return pm

    def forward(self, img):
        sp = self._nrv.spec
        assert tuple(img.shape[-2:]) == tuple(sp["image_size"]), \
            "expected images of size %s, got %s" % (sp["image_size"], tuple(img.shape[-2:]))
        assert sp["dim"] % 4 == 0, "feature dimension must be multiple of 4 for sincos emb"
        return _engine.run_model(self._nrv, img, with_head=True)
