"""VisionTransformer (torchvision-style ViT with the `robust` flag) — drop-in for the reference's
vit_pytorch_robust/vit.py:178-519.

Constructor arguments, factories (vit_b_16 ... vit_h_14), attribute tree and state_dict keys are
the reference's (which are torchvision's: class_token, conv_proj, encoder.pos_embedding,
encoder.layers.encoder_layer_i.{ln_1,self_attention,ln_2,mlp}, encoder.ln, heads.head), so
checkpoints, `model.heads.head = nn.Identity()` (examples/evaluation.py:129-131) and optimisers
keep working.  forward() runs the fused libnrvit encoder; the sub-modules only own parameters.

Note: the reference's own forward raises as shipped (utils.py:877 `asdf`; utils.py:210); the
semantics implemented here are those of the torchvision class it was copied from, with
robust=True meaning SinkhornAttention(-1, 3 iterations) (utils.py:1025-1037).
"""
import math
from collections import OrderedDict
from functools import partial
from typing import Any, Callable, Optional

import torch
import torch.nn as nn

from . import engine as _engine
from .simple_vit import _FusedOnly, Rearrange, Softmax, introspection_plan, pair

__all__ = ["VisionTransformer", "vit_b_16", "vit_b_32", "vit_l_16", "vit_l_32", "vit_h_14",
           "interpolate_embeddings", "ViT", "Transformer", "Attention", "FeedForward"]


def _check_dropout(p, what):
    """Dropout runs inside the fused encoder (nrv_dropout; include/nrvit.h NRV_DROP_*), see engine.dropout_request()."""
    if not 0.0 <= float(p) < 1.0:
        raise ValueError("%s must be in [0, 1), got %r" % (what, p))


class MLPBlock(nn.Sequential):
    """vit.py:35-84 — Linear, GELU, Dropout, Linear, Dropout with keys `0.*` and `3.*`."""

    _version = 2

    def __init__(self, in_dim: int, mlp_dim: int, dropout: float):
        super().__init__(nn.Linear(in_dim, mlp_dim), nn.GELU(), nn.Dropout(dropout),
                         nn.Linear(mlp_dim, in_dim), nn.Dropout(dropout))
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.normal_(m.bias, std=1e-6)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                              unexpected_keys, error_msgs):
        # legacy torchvision checkpoints name the two Linears linear_1 / linear_2 (vit.py:66-74)
        version = local_metadata.get("version", None)
        if version is None or version < 2:
            for old, new in (("linear_1", "0"), ("linear_2", "3")):
                for kind in ("weight", "bias"):
                    k = "%s%s.%s" % (prefix, old, kind)
                    if k in state_dict:
                        state_dict["%s%s.%s" % (prefix, new, kind)] = state_dict.pop(k)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys,
                                      unexpected_keys, error_msgs)

    def forward(self, x):
        raise NotImplementedError("MLPBlock only holds parameters: it runs inside the fused encoder")


class MultiheadAttention(_FusedOnly):
    """Parameter holder with nn.MultiheadAttention's packed layout (utils.py:600-728):
    in_proj_weight [3E, E] (q|k|v), in_proj_bias [3E], out_proj Linear(E, E)."""

    def __init__(self, embed_dim, num_heads, dropout=0.0, batch_first=True, robust=False):
        super().__init__()
        assert embed_dim % num_heads == 0, "embed_dim must be divisible by num_heads"
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.dropout, self.batch_first, self.robust = dropout, batch_first, robust
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.empty(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)     # utils.py:718-728
        nn.init.constant_(self.in_proj_bias, 0.0)
        nn.init.constant_(self.out_proj.bias, 0.0)


class EncoderBlock(_FusedOnly):
    """vit.py:87-130"""

    def __init__(self, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout,
                 norm_layer=partial(nn.LayerNorm, eps=1e-6), robust=False):
        super().__init__()
        self.num_heads = num_heads
        self.ln_1 = norm_layer(hidden_dim)
        self.self_attention = MultiheadAttention(hidden_dim, num_heads, dropout=attention_dropout,
                                                 batch_first=True, robust=robust)
        self.dropout = nn.Dropout(dropout)
        self.ln_2 = norm_layer(hidden_dim)
        self.mlp = MLPBlock(hidden_dim, mlp_dim, dropout)


class _FusedSequential(nn.Sequential):
    """`encoder.layers` (vit.py:160-167): an nn.Sequential by name and keys; executed only as part of the fused encoder."""
    _nrv_out = None

    def forward(self, x):
        if self._nrv_out is not None:
            return self._nrv_out
        raise NotImplementedError("encoder.layers only holds the EncoderBlocks: they run inside the fused encoder")


class Encoder(_FusedOnly):
    """vit.py:133-175"""

    def __init__(self, seq_length, num_layers, num_heads, hidden_dim, mlp_dim, dropout, attention_dropout,
                 norm_layer=partial(nn.LayerNorm, eps=1e-6), robust=False):
        super().__init__()
        self.pos_embedding = nn.Parameter(torch.empty(1, seq_length, hidden_dim).normal_(std=0.02))
        self.dropout = nn.Dropout(dropout)
        layers = OrderedDict()
        for i in range(num_layers):
            layers["encoder_layer_%d" % i] = EncoderBlock(num_heads, hidden_dim, mlp_dim, dropout,
                                                          attention_dropout, norm_layer, robust=robust)
        self.layers = _FusedSequential(layers)
        self.ln = norm_layer(hidden_dim)


class VisionTransformer(nn.Module):
    """vit.py:178-351"""

    def __init__(self, image_size: int, patch_size: int, num_layers: int, num_heads: int, hidden_dim: int,
                 mlp_dim: int, dropout: float = 0.0, attention_dropout: float = 0.0, num_classes: int = 1000,
                 representation_size: Optional[int] = None,
                 norm_layer: Callable[..., nn.Module] = partial(nn.LayerNorm, eps=1e-6),
                 conv_stem_configs=None, robust: bool = False):
        super().__init__()
        torch._assert(image_size % patch_size == 0, "Input shape indivisible by patch size!")
        if conv_stem_configs is not None:
            raise NotImplementedError("conv_stem_configs (vit.py:211-236) is outside the fused hot path")
        _check_dropout(dropout, "dropout")
        _check_dropout(attention_dropout, "attention_dropout")
        self.image_size, self.patch_size = image_size, patch_size
        self.hidden_dim, self.mlp_dim = hidden_dim, mlp_dim
        self.attention_dropout, self.dropout = attention_dropout, dropout
        self.num_classes, self.representation_size = num_classes, representation_size
        self.norm_layer, self.robust = norm_layer, robust

        self.conv_proj = nn.Conv2d(in_channels=3, out_channels=hidden_dim, kernel_size=patch_size, stride=patch_size)
        seq_length = (image_size // patch_size) ** 2
        self.class_token = nn.Parameter(torch.zeros(1, 1, hidden_dim))
        seq_length += 1
        self.encoder = Encoder(seq_length, num_layers, num_heads, hidden_dim, mlp_dim, dropout,
                               attention_dropout, norm_layer, robust=robust)
        self.seq_length = seq_length

        heads_layers = OrderedDict()
        if representation_size is None:
            heads_layers["head"] = nn.Linear(hidden_dim, num_classes)
        else:
            heads_layers["pre_logits"] = nn.Linear(hidden_dim, representation_size)
            heads_layers["act"] = nn.Tanh()
            heads_layers["head"] = nn.Linear(representation_size, num_classes)
        self.heads = nn.Sequential(heads_layers)

        # initialisation as vit.py:273-306
        fan_in = self.conv_proj.in_channels * self.conv_proj.kernel_size[0] * self.conv_proj.kernel_size[1]
        nn.init.trunc_normal_(self.conv_proj.weight, std=math.sqrt(1 / fan_in))
        nn.init.zeros_(self.conv_proj.bias)
        if hasattr(self.heads, "pre_logits"):
            fan_in = self.heads.pre_logits.in_features
            nn.init.trunc_normal_(self.heads.pre_logits.weight, std=math.sqrt(1 / fan_in))
            nn.init.zeros_(self.heads.pre_logits.bias)
        nn.init.zeros_(self.heads.head.weight)
        nn.init.zeros_(self.heads.head.bias)

        eps = getattr(self.encoder.ln, "eps", 1e-6)
        self._nrv = _engine.Engine(
            dict(image_size=(image_size, image_size), patch_size=(patch_size, patch_size), channels=3,
                 dim=hidden_dim, depth=num_layers, heads=num_heads, dim_head=hidden_dim // num_heads,
                 mlp_dim=mlp_dim, cls_token=True, pool="cls", patch_order="cp1p2", qkv_bias=True,
                 ln_eps=eps, robust=robust),
            self._nrv_param_map)

    # ---- engine plumbing ------------------------------------------------------------------
    def _fusable_head(self):
        """The classifier is fused (nrv_gemm) only while `heads` is still the plain Linear the
        constructor made; user-replaced heads (nn.Identity, probes) run as ordinary modules."""
        mods = list(self.heads.children())
        return len(mods) == 1 and isinstance(mods[0], nn.Linear) and mods[0].in_features == self.hidden_dim

    def _nrv_param_map(self):
        pm = {
            "w_patch": self.conv_proj.weight, "b_patch": self.conv_proj.bias,
            "pos": self.encoder.pos_embedding, "cls": self.class_token,
            "lnf_g": self.encoder.ln.weight, "lnf_b": self.encoder.ln.bias,
        }
        if self._fusable_head():
            head = list(self.heads.children())[0]
            pm["head_w"], pm["head_b"] = head.weight, head.bias
        for i, blk in enumerate(self.encoder.layers):
            pre = "l%d." % i
            pm[pre + "ln1_g"], pm[pre + "ln1_b"] = blk.ln_1.weight, blk.ln_1.bias
            pm[pre + "w_qkv"], pm[pre + "b_qkv"] = blk.self_attention.in_proj_weight, blk.self_attention.in_proj_bias
            pm[pre + "w_out"], pm[pre + "b_out"] = blk.self_attention.out_proj.weight, blk.self_attention.out_proj.bias
            pm[pre + "ln2_g"], pm[pre + "ln2_b"] = blk.ln_2.weight, blk.ln_2.bias
            pm[pre + "w_fc1"], pm[pre + "b_fc1"] = blk.mlp[0].weight, blk.mlp[0].bias
            pm[pre + "w_fc2"], pm[pre + "b_fc2"] = blk.mlp[3].weight, blk.mlp[3].bias
        return pm

    def forward(self, x: torch.Tensor):
        # vit.py:166,174 (embedding), :109,125 (after attention), :45,47 (MLP) share `dropout`; :105 attention_dropout
        drop = _engine.dropout_request(self.training, p=self.dropout, p_emb=self.dropout, p_attn=self.attention_dropout,
                                       robust=self.robust)
        n, c, h, w = x.shape
        torch._assert(h == self.image_size, f"Wrong image height! Expected {self.image_size} but got {h}!")
        torch._assert(w == self.image_size, f"Wrong image width! Expected {self.image_size} but got {w}!")
        plan = self._introspection_plan()
        if self._fusable_head():
            return _engine.run_model(self._nrv, x, with_head=True, drop=drop, introspect=plan)
        feat = _engine.run_model(self._nrv, x, with_head=False, drop=drop, introspect=plan)
        return self.heads(feat.float())

    def _introspection_plan(self):
        """Forward hooks on `encoder.layers` or on one EncoderBlock see that module's token input / output ([B, N, D]; an
        EncoderBlock contains both residual additions, vit.py:118-130, so its output is a residual-stream state)."""
        blocks = list(self.encoder.layers)
        hooked = [m for m in self.modules() if m is not self and _engine.has_forward_hooks(m)]
        if not hooked:
            return None
        ok = {id(self.encoder.layers)} | {id(b) for b in blocks}
        inside = {id(m) for m in self.encoder.modules()} | {id(self.conv_proj)}
        for m in hooked:
            if id(m) in inside and id(m) not in ok:
                raise NotImplementedError(
                    "forward hook on %s: inside the fused encoder only `encoder.layers` and its EncoderBlocks expose their "
                    "activations (there is no unfused fallback)" % type(m).__name__)

        def run(view):
            for l, b in enumerate(blocks):
                if _engine.has_forward_hooks(b):
                    _engine.emit(b, view.stream(2 * l), view.stream(2 * l + 2))
            if _engine.has_forward_hooks(self.encoder.layers):
                _engine.emit(self.encoder.layers, view.stream(0), view.stream(2 * view.L))
        return run


def _vision_transformer(patch_size, num_layers, num_heads, hidden_dim, mlp_dim, **kwargs: Any):
    image_size = kwargs.pop("image_size", 224)
    return VisionTransformer(image_size=image_size, patch_size=patch_size, num_layers=num_layers,
                             num_heads=num_heads, hidden_dim=hidden_dim, mlp_dim=mlp_dim, **kwargs)


def vit_b_16(**kwargs: Any) -> VisionTransformer:
    """vit.py:377-403"""
    return _vision_transformer(patch_size=16, num_layers=12, num_heads=12, hidden_dim=768, mlp_dim=3072, **kwargs)


def vit_b_32(**kwargs: Any) -> VisionTransformer:
    """vit.py:406-432"""
    return _vision_transformer(patch_size=32, num_layers=12, num_heads=12, hidden_dim=768, mlp_dim=3072, **kwargs)


def vit_l_16(**kwargs: Any) -> VisionTransformer:
    """vit.py:435-461"""
    return _vision_transformer(patch_size=16, num_layers=24, num_heads=16, hidden_dim=1024, mlp_dim=4096, **kwargs)


def vit_l_32(**kwargs: Any) -> VisionTransformer:
    """vit.py:464-490"""
    return _vision_transformer(patch_size=32, num_layers=24, num_heads=16, hidden_dim=1024, mlp_dim=4096, **kwargs)


def vit_h_14(**kwargs: Any) -> VisionTransformer:
    """vit.py:493-519"""
    return _vision_transformer(patch_size=14, num_layers=32, num_heads=16, hidden_dim=1280, mlp_dim=5120, **kwargs)


def interpolate_embeddings(image_size, patch_size, model_state, interpolation_mode="bicubic", reset_heads=False):
    """vit.py:522-603 is a verbatim copy of torchvision's checkpoint helper; use torchvision's."""
    from torchvision.models.vision_transformer import interpolate_embeddings as _tv
    return _tv(image_size, patch_size, model_state, interpolation_mode, reset_heads)


# ------------------------------------------------------------------------------------------------
# README `ViT` (lucidrains API, README.md:67-111).  The reference's vit.py does not define it (which
# is why `from vit_pytorch_robust.vit import ViT` in distill.py:4 / mae.py:6 / recorder.py:5 fails);
# the structure below follows the in-tree statements of that API (learnable_memory_vit.py:30-151,
# vit_with_patch_dropout.py:54-152): LayerNorm inside Attention / FeedForward, packed to_qkv without
# bias, to_out = Sequential(Linear, Dropout), FeedForward.net = (LN, Linear, GELU, Dropout, Linear,
# Dropout), class token + learned pos_embedding [1, n+1, dim], pool in {cls, mean}, mlp_head = LN + Linear.
# PARITY UNPINNED by the reference (no runnable class); oracle = oracle/vit_oracle.py::readme_vit_forward.
# ------------------------------------------------------------------------------------------------
class FeedForward(_FusedOnly):
    def __init__(self, dim, hidden_dim, dropout=0.):
        super().__init__()
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(dropout))


class Attention(_FusedOnly):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.):
        super().__init__()
        inner_dim = dim_head * heads
        if heads == 1 and dim_head == dim:
            raise NotImplementedError("heads == 1 and dim_head == dim (to_out = Identity) is outside the fused path")
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.norm = nn.LayerNorm(dim)
        self.attend = Softmax(dim=-1)
        self.dropout = nn.Dropout(dropout)
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout))


class Transformer(_FusedOnly):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout),
                FeedForward(dim, mlp_dim, dropout=dropout),
            ]))


class ViT(nn.Module):
    def __init__(self, *, image_size, patch_size, num_classes, dim, depth, heads, mlp_dim, pool='cls', channels=3,
                 dim_head=64, dropout=0., emb_dropout=0.):
        super().__init__()
        image_height, image_width = pair(image_size)
        patch_height, patch_width = pair(patch_size)
        assert image_height % patch_height == 0 and image_width % patch_width == 0, \
            'Image dimensions must be divisible by the patch size.'
        num_patches = (image_height // patch_height) * (image_width // patch_width)
        patch_dim = channels * patch_height * patch_width
        assert pool in {'cls', 'mean'}, 'pool type must be either cls (cls token) or mean (mean pooling)'

        self.to_patch_embedding = nn.Sequential(
            Rearrange('b c (h p1) (w p2) -> b (h w) (p1 p2 c)', p1=patch_height, p2=patch_width),
            nn.Linear(patch_dim, dim),
        )
        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches + 1, dim))
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.dropout = nn.Dropout(emb_dropout)
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim, dropout)
        self.pool = pool
        self.to_latent = nn.Identity()
        self.mlp_head = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, num_classes))
        self._p_drop, self._p_emb = float(dropout), float(emb_dropout)

        self._nrv = _engine.Engine(
            dict(image_size=(image_height, image_width), patch_size=(patch_height, patch_width), channels=channels,
                 dim=dim, depth=depth, heads=heads, dim_head=dim_head, mlp_dim=mlp_dim, cls_token=True, pool=pool,
                 patch_order="p1p2c", qkv_bias=False, ln_eps=1e-5, robust=False),
            self._nrv_param_map)

    def _nrv_param_map(self):
        pm = {
            "w_patch": self.to_patch_embedding[1].weight, "b_patch": self.to_patch_embedding[1].bias,
            "pos": self.pos_embedding, "cls": self.cls_token,
            "lnf_g": self.mlp_head[0].weight, "lnf_b": self.mlp_head[0].bias,
            "head_w": self.mlp_head[1].weight, "head_b": self.mlp_head[1].bias,
        }
        for i, (attn, ff) in enumerate(self.transformer.layers):
            pre = "l%d." % i
            pm[pre + "ln1_g"], pm[pre + "ln1_b"] = attn.norm.weight, attn.norm.bias
            pm[pre + "w_qkv"] = attn.to_qkv.weight
            pm[pre + "w_out"], pm[pre + "b_out"] = attn.to_out[0].weight, attn.to_out[0].bias
            pm[pre + "ln2_g"], pm[pre + "ln2_b"] = ff.net[0].weight, ff.net[0].bias
            pm[pre + "w_fc1"], pm[pre + "b_fc1"] = ff.net[1].weight, ff.net[1].bias
            pm[pre + "w_fc2"], pm[pre + "b_fc2"] = ff.net[4].weight, ff.net[4].bias
        return pm

    def forward(self, img):
        # README ViT: `dropout` is used after softmax, after to_out and in the FeedForward; `emb_dropout` after pos
        drop = _engine.dropout_request(self.training, p=self._p_drop, p_emb=self._p_emb, p_attn=self._p_drop)
        sp = self._nrv.spec
        assert tuple(img.shape[-2:]) == tuple(sp["image_size"]), \
            "expected images of size %s, got %s" % (sp["image_size"], tuple(img.shape[-2:]))
        plan = introspection_plan(self, self.transformer, [attn.attend for attn, _ in self.transformer.layers])
        return _engine.run_model(self._nrv, img, with_head=True, drop=drop, introspect=plan)
