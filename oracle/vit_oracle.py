"""ORACLE — test infrastructure only.  CPU restatement of the reference's ViT encoder hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this file; the product (noise-robust-vit_b200/) never does.

Every function restates, in plain functional torch on the CPU (dtype-generic: float32 / float64),
what the cited reference code computes.  State is passed as a ``state_dict`` with the reference's
own keys, so reference weights drop straight in.  Backward = torch autograd over this restatement
(the reference's backward is autograd over the same graph).

Pinning (see tests/golden/make_golden.py, tests/test_oracle.py):
  * SimpleViT path — checked against the reference module itself (vit_pytorch_robust/simple_vit.py
    imported from /root/reference with the broken package __init__ bypassed), robust on and off:
    committed golden vectors in tests/golden/*.npz.
  * VisionTransformer path — the reference forward raises as shipped (utils.py:877 `asdf`,
    utils.py:210 permute on a 4-D tensor), so the oracle is pinned against the state-dict-identical
    torchvision.models.vision_transformer.VisionTransformer twin (the class vit.py:178-351 was
    copied from); robust=True is DEFINED as SinkhornAttention(-1, 3 iterations) (utils.py:1025-1037).
  * README `ViT` (lucidrains API) — `vit.py` defines no such class, but vit_with_patch_dropout.py:101-152 is a runnable
    in-tree class with exactly that structure.  With patch_dropout = 0 it equals the README model whose class-token
    row of pos_embedding is zero; readme_vit_forward is pinned against it (state-dict key map:
    readme_state_from_patch_dropout_vit) — golden fixture readme_vit.npz and a live test with seeded dropout masks on
    every site INCLUDING the attention probabilities (vit_with_patch_dropout.py:78-79).
  * Dropout (train mode) — masks are an INPUT of the restatement (`drop` callback).  The sites after the embedding,
    after attention and inside the MLP are pinned against the torchvision twin with shared seeded masks
    (tests/test_oracle.py); the attention-probability site is unpinned.
"""
import math

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# elementary pieces
# ------------------------------------------------------------------------------------------------
def posemb_sincos_2d(h, w, dim, temperature=10000, dtype=torch.float32):
    """simple_vit.py:15-28 — token t = y*w + x ; pe = [sin(x w), cos(x w), sin(y w), cos(y w)]."""
    assert dim % 4 == 0, "feature dimension must be multiple of 4 for sincos emb"
    y, x = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    omega = torch.arange(dim // 4) / (dim // 4 - 1)
    omega = 1.0 / (temperature ** omega)
    y = y.flatten()[:, None] * omega[None, :]
    x = x.flatten()[:, None] * omega[None, :]
    pe = torch.cat((x.sin(), x.cos(), y.sin(), y.cos()), dim=1)
    return pe.to(dtype)


def patchify_p1p2c(img, ph, pw):
    """einops 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)' (simple_vit.py:127-129): channel fastest."""
    b, c, H, W = img.shape
    h, w = H // ph, W // pw
    x = img.reshape(b, c, h, ph, w, pw).permute(0, 2, 4, 3, 5, 1)  # b h w p1 p2 c
    return x.reshape(b, h * w, ph * pw * c)


def patchify_cp1p2(img, ph, pw):
    """im2col of Conv2d(k=s=P) (vit.py:237-242,323-331): flattening order (c p1 p2)."""
    b, c, H, W = img.shape
    h, w = H // ph, W // pw
    x = img.reshape(b, c, h, ph, w, pw).permute(0, 2, 4, 1, 3, 5)  # b h w c p1 p2
    return x.reshape(b, h * w, c * ph * pw)


def layer_norm(x, g, b, eps):
    """nn.LayerNorm over the last dim, biased variance (simple_vit.py:38,54,136 ; vit.py:104,115,167)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def gelu_erf(x):
    """nn.GELU() default = exact erf form (simple_vit.py:40 ; vit.py:44)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def sinkhorn3(p):
    """utils.py:1031-1037 applied to softmax output: 3 x (row-normalise, column-normalise), then
    one more row normalisation; no epsilon."""
    for _ in range(3):
        p = p / p.sum(-1, keepdim=True)
        p = p / p.sum(-2, keepdim=True)
    return p / p.sum(-1, keepdim=True)


def attention_core(q, k, v, scale, robust, drop=None):
    """simple_vit.py:70-74 (and the intended utils.py:207-232): softmax(q k^T * scale) v per head.
    q,k,v: [B,H,N,dh].  drop: optional callable on the probabilities [B,H,N,N] (dropout_p of
    nn.MultiheadAttention, utils.py:225-228 ; README ViT Attention.dropout)."""
    dots = torch.matmul(q, k.transpose(-1, -2)) * scale
    attn = torch.softmax(dots, dim=-1)
    if robust:
        attn = sinkhorn3(attn)
    if drop is not None:
        attn = drop(attn)
    return torch.matmul(attn, v), attn


def split_heads(t, heads):
    """'b n (h d) -> b h n d' (simple_vit.py:68 ; utils.py:489-502,568-570)."""
    b, n, hd = t.shape
    return t.reshape(b, n, heads, hd // heads).permute(0, 2, 1, 3)


def merge_heads(t):
    b, h, n, d = t.shape
    return t.permute(0, 2, 1, 3).reshape(b, n, h * d)


def cross_entropy(logits, labels, label_smoothing=0.0):
    """F.cross_entropy(preds, y, label_smoothing) mean reduction (examples/baseline.py:70)."""
    logp = torch.log_softmax(logits, dim=-1)
    nll = -logp.gather(1, labels[:, None]).squeeze(1)
    smooth = -logp.mean(dim=-1)
    return ((1.0 - label_smoothing) * nll + label_smoothing * smooth).mean()


def adamw_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-2):
    """torch.optim.AdamW single-tensor update (examples/CIFAR100.py:90-97); returns new (p, m, v)."""
    p = p * (1.0 - lr * weight_decay)
    m = beta1 * m + (1.0 - beta1) * g
    v = beta2 * v + (1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def clip_coef(total_norm, max_norm):
    """torch.nn.utils.clip_grad_norm_ coefficient (grad_max_norm, examples/CIFAR100.py:192)."""
    return min(1.0, max_norm / (total_norm + 1e-6))


# ------------------------------------------------------------------------------------------------
# SimpleViT  (simple_vit.py:100-149)
# ------------------------------------------------------------------------------------------------
def simple_vit_forward(sd, img, *, patch_size, heads, dim_head=64, robust=False, return_tokens=False, return_attn=False):
    """sd: state_dict with the reference keys (to_patch_embedding.1.*, transformer.layers.i.{0,1}.*,
    linear_head.{0,1}.*).  img [B,C,H,W].  Returns logits [B,num_classes]; return_tokens: the transformer's output
    tokens instead (what extractor.py:50-59 hooks); return_attn: (logits, [per-layer `attend` outputs [B,H,N,N]]), what
    recorder.py:28-31 records."""
    ph, pw = (patch_size, patch_size) if isinstance(patch_size, int) else patch_size
    dt = img.dtype
    P = lambda k: sd[k].to(dt)  # noqa: E731
    B, C, H, W = img.shape
    h, w = H // ph, W // pw
    x = patchify_p1p2c(img, ph, pw)                                                  # :127-129
    x = x @ P("to_patch_embedding.1.weight").t() + P("to_patch_embedding.1.bias")    # :130
    dim = x.shape[-1]
    x = x + posemb_sincos_2d(h, w, dim).to(dt)                                       # :141-143
    depth = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.layers."))
    scale = dim_head ** -0.5                                                         # :53
    attns = []
    for i in range(depth):
        pa = "transformer.layers.%d.0." % i
        pf = "transformer.layers.%d.1." % i
        y = layer_norm(x, P(pa + "norm.weight"), P(pa + "norm.bias"), 1e-5)          # :65
        qkv = y @ P(pa + "to_qkv.weight").t()                                        # :67 (no bias)
        q, k, v = (split_heads(t, heads) for t in qkv.chunk(3, dim=-1))              # :67-68
        o, attn = attention_core(q, k, v, scale, robust)                             # :70-74
        attns.append(attn)
        x = merge_heads(o) @ P(pa + "to_out.weight").t() + x                         # :75-76,95
        y = layer_norm(x, P(pf + "net.0.weight"), P(pf + "net.0.bias"), 1e-5)        # :38
        y = gelu_erf(y @ P(pf + "net.1.weight").t() + P(pf + "net.1.bias"))          # :39-40
        x = y @ P(pf + "net.3.weight").t() + P(pf + "net.3.bias") + x                # :41,96
    if return_tokens:
        return x
    x = x.mean(dim=1)                                                                # :146
    x = layer_norm(x, P("linear_head.0.weight"), P("linear_head.0.bias"), 1e-5)      # :136
    logits = x @ P("linear_head.1.weight").t() + P("linear_head.1.bias")
    return (logits, attns) if return_attn else logits


# ------------------------------------------------------------------------------------------------
# VisionTransformer  (vit.py:178-351; attention semantics = nn.MultiheadAttention(batch_first=True))
# ------------------------------------------------------------------------------------------------
# dropout sites, numbered as NRV_DROP_* in include/nrvit.h
DROP_ATTN_OUT, DROP_FC1, DROP_FC2, DROP_EMB, DROP_ATTN_PROB = 0, 1, 2, 3, 4


def vision_transformer_forward(sd, img, *, patch_size, num_heads, robust=False, eps=1e-6,
                               return_features=False, drop=None, return_streams=False):
    """sd: state_dict with torchvision/reference keys (class_token, conv_proj.*, encoder.*, heads.*).
    drop: None (eval / p = 0) or a callable drop(x, layer, site) -> x * mask / (1 - p) standing for the
    nn.Dropout modules of a train()-mode forward (vit.py:45,47 MLP ; :125 after attention ; :174 embedding;
    layer = -1 for the embedding).  The masks are an input of the restatement, so that a test can hand it the
    masks the implementation under test drew.  return_streams: (logits, [tokens entering block 0, leaving block 0, ...]),
    the input / output of every EncoderBlock (vit.py:118-130)."""
    if drop is None:
        drop = lambda t, layer, site: t  # noqa: E731
    dt = img.dtype
    P = lambda k: sd[k].to(dt)  # noqa: E731
    B = img.shape[0]
    D = sd["class_token"].shape[-1]
    p = patch_size
    x = patchify_cp1p2(img, p, p)                                                    # vit.py:323-331
    x = x @ P("conv_proj.weight").reshape(D, -1).t() + P("conv_proj.bias")           # vit.py:237-242
    x = torch.cat([P("class_token").expand(B, -1, -1), x], dim=1)                    # vit.py:341-342
    x = drop(x + P("encoder.pos_embedding"), -1, DROP_EMB)                           # vit.py:174
    depth = 1 + max(int(k.split("encoder_layer_")[1].split(".")[0]) for k in sd if "encoder_layer_" in k)
    dh = D // num_heads
    streams = [x]
    for i in range(depth):
        pre = "encoder.layers.encoder_layer_%d." % i
        y = layer_norm(x, P(pre + "ln_1.weight"), P(pre + "ln_1.bias"), eps)         # vit.py:123
        qkv = y @ P(pre + "self_attention.in_proj_weight").t() + P(pre + "self_attention.in_proj_bias")  # utils.py:415
        q, k, v = (split_heads(t, num_heads) for t in qkv.chunk(3, dim=-1))
        o, _ = attention_core(q, k, v, 1.0 / math.sqrt(dh), robust,                  # utils.py:212-213,225-228
                              (lambda a, i=i: drop(a, i, DROP_ATTN_PROB)))
        o = merge_heads(o) @ P(pre + "self_attention.out_proj.weight").t() + P(pre + "self_attention.out_proj.bias")
        x = drop(o, i, DROP_ATTN_OUT) + x                                            # vit.py:125-126
        y = layer_norm(x, P(pre + "ln_2.weight"), P(pre + "ln_2.bias"), eps)         # vit.py:128
        y = gelu_erf(y @ P(pre + "mlp.0.weight").t() + P(pre + "mlp.0.bias"))        # vit.py:41-44
        y = drop(y, i, DROP_FC1)                                                     # vit.py:45
        x = drop(y @ P(pre + "mlp.3.weight").t() + P(pre + "mlp.3.bias"), i, DROP_FC2) + x   # vit.py:46-47,129-130
        streams.append(x)
    if return_streams:
        return vision_transformer_forward(sd, img, patch_size=patch_size, num_heads=num_heads, robust=robust, eps=eps,
                                          return_features=return_features, drop=drop), streams
    x = layer_norm(x, P("encoder.ln.weight"), P("encoder.ln.bias"), eps)             # vit.py:175
    x = x[:, 0]                                                                      # vit.py:347
    if return_features or "heads.head.weight" not in sd:
        return x
    if "heads.pre_logits.weight" in sd:                                              # vit.py:266-269
        x = torch.tanh(x @ P("heads.pre_logits.weight").t() + P("heads.pre_logits.bias"))
    return x @ P("heads.head.weight").t() + P("heads.head.bias")                     # vit.py:265,349


# ------------------------------------------------------------------------------------------------
# README ViT (lucidrains API), restated from vit_with_patch_dropout.py:54-152 and README.md:67-111;
# pinned against the reference class vit_with_patch_dropout.ViT(patch_dropout=0) through the key map below
# ------------------------------------------------------------------------------------------------
def readme_state_from_patch_dropout_vit(sd):
    """state_dict of the reference's vit_with_patch_dropout.ViT -> keys / shapes of the README `ViT`:
    PreNorm(norm, fn) wrappers become the LayerNorm-inside modules of the README class (`.fn.` dropped, FeedForward's
    norm becomes net.0 and its Linears shift to net.1 / net.4); pos_embedding [n, D] (added before the class token is
    concatenated, vit_with_patch_dropout.py:136-143) becomes [1, n+1, D] with a zero class-token row."""
    out = {}
    for k, v in sd.items():
        if k == "pos_embedding":
            out[k] = torch.cat([torch.zeros_like(v[:1]), v], dim=0)[None]
            continue
        parts = k.split(".")
        if parts[0] == "transformer":
            i, which = parts[2], parts[3]
            rest = parts[4:]
            if which == "0":
                rest = rest[1:] if rest[0] == "fn" else rest
            else:
                if rest[0] == "norm":
                    rest = ["net", "0"] + rest[1:]
                else:  # fn.net.{0,3}
                    rest = ["net", {"0": "1", "3": "4"}[rest[2]]] + rest[3:]
            k = ".".join(parts[:3] + [which] + rest)
        out[k] = v
    return out


def readme_vit_forward(sd, img, *, patch_size, heads, dim_head=64, pool="cls", drop=None):
    """drop(x, layer, site): see vision_transformer_forward (README `dropout`: sites ATTN_PROB, ATTN_OUT, FC1, FC2;
    `emb_dropout`: site EMB)."""
    if drop is None:
        drop = lambda t, layer, site: t  # noqa: E731
    ph, pw = (patch_size, patch_size) if isinstance(patch_size, int) else patch_size
    dt = img.dtype
    P = lambda k: sd[k].to(dt)  # noqa: E731
    B = img.shape[0]
    x = patchify_p1p2c(img, ph, pw)
    x = x @ P("to_patch_embedding.1.weight").t() + P("to_patch_embedding.1.bias")
    x = torch.cat([P("cls_token").expand(B, -1, -1), x], dim=1)
    x = drop(x + P("pos_embedding")[:, : x.shape[1]], -1, DROP_EMB)
    depth = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.layers."))
    scale = dim_head ** -0.5
    for i in range(depth):
        pa = "transformer.layers.%d.0." % i
        pf = "transformer.layers.%d.1." % i
        y = layer_norm(x, P(pa + "norm.weight"), P(pa + "norm.bias"), 1e-5)
        qkv = y @ P(pa + "to_qkv.weight").t()
        q, k, v = (split_heads(t, heads) for t in qkv.chunk(3, dim=-1))
        o, _ = attention_core(q, k, v, scale, False, (lambda a, i=i: drop(a, i, DROP_ATTN_PROB)))
        o = merge_heads(o)
        if (pa + "to_out.0.weight") in sd:
            o = drop(o @ P(pa + "to_out.0.weight").t() + P(pa + "to_out.0.bias"), i, DROP_ATTN_OUT)
        x = o + x
        y = layer_norm(x, P(pf + "net.0.weight"), P(pf + "net.0.bias"), 1e-5)
        y = drop(gelu_erf(y @ P(pf + "net.1.weight").t() + P(pf + "net.1.bias")), i, DROP_FC1)
        x = drop(y @ P(pf + "net.4.weight").t() + P(pf + "net.4.bias"), i, DROP_FC2) + x
    x = x.mean(dim=1) if pool == "mean" else x[:, 0]
    x = layer_norm(x, P("mlp_head.0.weight"), P("mlp_head.0.bias"), 1e-5)
    return x @ P("mlp_head.1.weight").t() + P("mlp_head.1.bias")


# ------------------------------------------------------------------------------------------------
# loss + gradients helper
# ------------------------------------------------------------------------------------------------
def loss_and_grads(forward_fn, sd, img, labels, label_smoothing=0.0, **kw):
    """Runs forward_fn(sd, img, **kw) with autograd on every floating tensor of sd; returns
    (logits, loss, {key: grad})."""
    leaf = {k: v.detach().clone().to(img.dtype).requires_grad_(True) for k, v in sd.items()
            if v.is_floating_point()}
    logits = forward_fn(leaf, img, **kw)
    loss = cross_entropy(logits, labels, label_smoothing)
    grads = torch.autograd.grad(loss, list(leaf.values()), allow_unused=True)
    return logits.detach(), loss.detach(), {k: g for k, g in zip(leaf.keys(), grads) if g is not None}


def rel_l2(a, b):
    a = a.detach().double().flatten().cpu()
    b = b.detach().double().flatten().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def cosine(a, b):
    a = a.detach().double().flatten().cpu()
    b = b.detach().double().flatten().cpu()
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)).item()
