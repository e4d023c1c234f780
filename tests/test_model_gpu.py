"""End-to-end parity of the product modules (CUDA, through libnrvit) against the oracle / golden
vectors: logits and every parameter gradient.

Tolerances (BASELINE.json north_star): rel-L2 <= 1e-3 in the FP32 check mode, cosine >= 0.999 in
BF16 mode.  The check mode is far tighter in practice (3xTF32 split GEMMs); the asserted bound for
it is 2e-4, the BF16 bound is the stated 0.999 cosine plus rel-L2 <= 3e-2 on the logits."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

import vit_oracle as O  # noqa: E402
import vit_pytorch_robust as V  # noqa: E402
from helpers import (README_CFG, SIMPLE_CFG, VIT_CFG, compare_grads, load_golden, load_readme_golden,  # noqa: E402
                     model_loss_and_grads, randomize_)

DEV = "cuda:0"
CHECK_REL = 2e-4
BF16_COS = 0.999


def set_mode(model, dtype):
    model._nrv.compute_dtype = dtype


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("robust", [False, True], ids=["softmax", "robust"])
def test_simplevit_matches_golden(golden_dir, dtype, robust):
    """Golden vectors come from the reference module itself (robust=True = SinkhornAttention)."""
    name = "simplevit_robust.npz" if robust else "simplevit_softmax.npz"
    sd, grads, img, labels, logits, loss = load_golden(os.path.join(golden_dir, name))
    m = V.SimpleViT(**SIMPLE_CFG, robust=robust)
    m.load_state_dict(sd)
    m = m.to(DEV)
    set_mode(m, dtype)
    lg, ls, gr = model_loss_and_grads(m, img.to(DEV), labels.to(DEV), 0.1)
    if dtype == torch.float32:
        assert O.rel_l2(lg, logits) < CHECK_REL
        assert abs(ls - loss) < 1e-4
        worst, key = compare_grads(gr, grads, O.rel_l2)
        assert worst < CHECK_REL, (key, worst)
    else:
        assert O.cosine(lg, logits) > BF16_COS
        assert O.rel_l2(lg, logits) < 3e-2
        worst, key = compare_grads(gr, grads, O.cosine)
        assert worst > BF16_COS, (key, worst)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_visiontransformer_matches_golden(golden_dir, dtype):
    sd, grads, img, labels, logits, loss = load_golden(os.path.join(golden_dir, "visiontransformer_softmax.npz"))
    m = V.VisionTransformer(**VIT_CFG)
    m.load_state_dict(sd)
    m = m.to(DEV)
    set_mode(m, dtype)
    lg, ls, gr = model_loss_and_grads(m, img.to(DEV), labels.to(DEV), 0.1)
    assert set(gr) == set(grads)
    if dtype == torch.float32:
        assert O.rel_l2(lg, logits) < CHECK_REL
        worst, key = compare_grads(gr, grads, O.rel_l2)
        assert worst < CHECK_REL, (key, worst)
    else:
        assert O.cosine(lg, logits) > BF16_COS
        worst, key = compare_grads(gr, grads, O.cosine)
        assert worst > BF16_COS, (key, worst)


CASES = [
    # name, kind, ctor kwargs, batch
    ("simple_rect", "simple", dict(image_size=(24, 40), patch_size=(8, 4), num_classes=7, dim=48, depth=3, heads=3,
                                   mlp_dim=96, dim_head=16), 3),
    ("simple_cifar", "simple", dict(image_size=32, patch_size=4, num_classes=100, dim=128, depth=2, heads=4,
                                    mlp_dim=256, dim_head=32), 5),
    ("simple_readme_small", "simple", dict(image_size=64, patch_size=32, num_classes=1000, dim=256, depth=1, heads=4,
                                           mlp_dim=512), 2),
    ("vit_p16", "vit", dict(image_size=48, patch_size=16, num_layers=2, num_heads=4, hidden_dim=128, mlp_dim=256,
                            num_classes=16), 4),
    ("vit_p14_ragged_patchdim", "vit", dict(image_size=28, patch_size=14, num_layers=1, num_heads=2, hidden_dim=160,
                                            mlp_dim=320, num_classes=24), 2),
    # dh = 64 with 197 and 226 tokens: the tcgen05 attention kernels (two key tiles, ragged second tile) and the bias
    # gradients fused into the attention-backward and FC2-dX epilogues
    ("vit_197_tokens_dh64", "vit", dict(image_size=224, patch_size=16, num_layers=1, num_heads=2, hidden_dim=128,
                                        mlp_dim=256, num_classes=12), 2),
    ("vit_226_tokens_dh64", "vit", dict(image_size=240, patch_size=16, num_layers=1, num_heads=2, hidden_dim=128,
                                        mlp_dim=256, num_classes=12), 2),
]


@pytest.mark.parametrize("ln_mode", ["separate", "folded"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_models_match_oracle(case, dtype, ln_mode):
    """Logits and every parameter gradient against the oracle, with the LayerNorms in front of the QKV / FC1 projections as
    stand-alone kernels and folded into those GEMMs (row statistics from the producing GEMM's epilogue; the backward pass
    gets the normalised rows from the LayerNorm backward kernel)."""
    from vit_pytorch_robust import _abi
    name, kind, kw, B = case
    torch.manual_seed(0)
    m = V.SimpleViT(**kw) if kind == "simple" else V.VisionTransformer(**kw)
    randomize_(m, seed=hash(name) % 1000)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    size = kw["image_size"] if isinstance(kw["image_size"], tuple) else (kw["image_size"],) * 2
    g = torch.Generator().manual_seed(1)
    img = torch.randn(B, 3, *size, generator=g)
    ncls = kw["num_classes"]
    labels = torch.randint(0, ncls, (B,), generator=g)
    if kind == "simple":
        fwd = lambda s, x: O.simple_vit_forward(s, x, patch_size=kw["patch_size"], heads=kw["heads"],  # noqa: E731
                                                dim_head=kw.get("dim_head", 64))
    else:
        fwd = lambda s, x: O.vision_transformer_forward(s, x, patch_size=kw["patch_size"],  # noqa: E731
                                                        num_heads=kw["num_heads"])
    ref_logits, ref_loss, ref_grads = O.loss_and_grads(lambda s, x: fwd(s, x), sd, img.double(), labels, 0.1)
    m = m.to(DEV)
    set_mode(m, dtype)
    m._nrv.ln_mode_train = _abi.LN_FOLDED if ln_mode == "folded" else _abi.LN_SEPARATE
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
    # inference (eval + no_grad): both LayerNorm modes, with and without the CUDA-graph replay of small batches
    m.eval()
    m._nrv.ln_fold_min_tokens = 0
    for infer_mode in (_abi.LN_FOLDED, _abi.LN_SEPARATE):
        m._nrv.ln_mode_infer = infer_mode
        m._nrv._graphs.clear()
        with torch.no_grad():
            a = m(img.to(DEV)).float().cpu()
            b = m(img.to(DEV)).float().cpu()          # second call replays the captured graph
        assert torch.equal(a, b)
        if dtype == torch.float32:
            assert O.rel_l2(a, ref_logits) < CHECK_REL
        else:
            assert O.cosine(a, ref_logits) > BF16_COS


def test_inference_matches_training_forward_and_eval_mode():
    m = V.VisionTransformer(**VIT_CFG)
    randomize_(m, 3)
    m = m.to(DEV)
    m._nrv.ln_mode_train = m._nrv.ln_mode_infer        # same LayerNorm mode (folded) on both paths: bit-identical results
    m._nrv.ln_fold_min_tokens = 0
    x = torch.randn(3, 3, 32, 32, device=DEV)
    a = m(x)
    m.eval()
    with torch.no_grad():
        b = m(x)
    assert torch.equal(a.detach(), b)
    assert not b.requires_grad
    # under no_grad the forward must take the inference path (buffer rotation, no activation stash) although every
    # parameter still requires grad: buffer keys are (batch, training, ...)
    assert [k[1] for k in m._nrv._bufs] == [1] and len(m._nrv._graphs) == 1   # training buffers + one captured inference graph


def test_replaced_head_and_frozen_backbone():
    """examples/evaluation.py:129-140 — heads.head = Identity, requires_grad_(False), eval()."""
    m = V.VisionTransformer(**VIT_CFG)
    randomize_(m, 4)
    m.heads.head = torch.nn.Identity()
    m = m.to(DEV)
    m.requires_grad_(False)
    m.eval()
    x = torch.randn(2, 3, 32, 32, device=DEV)
    f = m(x)
    assert f.shape == (2, 64) and f.dtype == torch.float32
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ref = O.vision_transformer_forward(sd, x.cpu().double(), patch_size=8, num_heads=2, return_features=True)
    assert O.cosine(f, ref) > BF16_COS


def test_gradient_accumulation_and_zero_grad():
    m = V.SimpleViT(**SIMPLE_CFG).to(DEV)
    m._nrv.compute_dtype = torch.float32
    x = torch.randn(2, 3, 32, 32, device=DEV)
    y = torch.randint(0, 10, (2,), device=DEV)
    torch.nn.functional.cross_entropy(m(x), y).backward()
    g1 = {k: p.grad.clone() for k, p in m.named_parameters()}
    torch.nn.functional.cross_entropy(m(x), y).backward()
    for k, p in m.named_parameters():
        assert O.rel_l2(p.grad, 2 * g1[k]) < 1e-5, k
    m.zero_grad(set_to_none=True)
    torch.nn.functional.cross_entropy(m(x), y).backward()
    for k, p in m.named_parameters():
        assert O.rel_l2(p.grad, g1[k]) < 1e-5, k


def test_training_step_with_fused_adamw_tracks_torch_adamw():
    """Three optimiser steps: product (FusedAdamW, check mode) vs oracle graph + torch.optim.AdamW."""
    torch.manual_seed(0)
    m = V.SimpleViT(**SIMPLE_CFG)
    randomize_(m, 9)
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ref_params = {k: torch.nn.Parameter(v.double().clone()) for k, v in sd0.items()}
    ref_opt = torch.optim.AdamW(ref_params.values(), lr=1e-3, weight_decay=0.01)
    m = m.to(DEV)
    m._nrv.compute_dtype = torch.float32
    opt = V.FusedAdamW(m.parameters(), lr=1e-3, weight_decay=0.01)
    g = torch.Generator().manual_seed(2)
    for step in range(3):
        img = torch.randn(4, 3, 32, 32, generator=g)
        labels = torch.randint(0, 10, (4,), generator=g)
        opt.zero_grad()
        loss = V.softmax_cross_entropy(m(img.to(DEV)), labels.to(DEV), 0.1)
        loss.backward()
        opt.step()
        ref_opt.zero_grad()
        rl = O.cross_entropy(O.simple_vit_forward(ref_params, img.double(), patch_size=8, heads=2, dim_head=32), labels, 0.1)
        rl.backward()
        ref_opt.step()
        assert abs(loss.item() - rl.item()) < 1e-4
    for k, p in m.named_parameters():
        assert O.rel_l2(p.detach().cpu(), ref_params[k].detach()) < 1e-4, k


def test_bf16_shadow_follows_foreign_optimizer():
    m = V.SimpleViT(**SIMPLE_CFG).to(DEV)
    opt = torch.optim.SGD(m.parameters(), lr=0.5)
    x = torch.randn(2, 3, 32, 32, device=DEV)
    y = torch.randint(0, 10, (2,), device=DEV)
    a = m(x).detach().clone()
    torch.nn.functional.cross_entropy(m(x), y).backward()
    opt.step()
    b = m(x).detach()
    assert not torch.equal(a, b)  # parameters changed under us -> shadow was refreshed
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ref = O.simple_vit_forward(sd, x.cpu().double(), patch_size=8, heads=2, dim_head=32)
    assert O.cosine(b, ref) > BF16_COS


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_robust_visiontransformer_matches_oracle(dtype):
    """vit.VisionTransformer(robust=True): attention probabilities := SinkhornAttention(-1, 3 iterations)."""
    kw = dict(image_size=48, patch_size=8, num_layers=2, num_heads=2, hidden_dim=128, mlp_dim=256, num_classes=16)
    m = V.VisionTransformer(**kw, robust=True)
    randomize_(m, 21)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    img = torch.randn(3, 3, 48, 48, generator=g)
    labels = torch.randint(0, 16, (3,), generator=g)
    ref_logits, _, ref_grads = O.loss_and_grads(
        lambda s_, x: O.vision_transformer_forward(s_, x, patch_size=8, num_heads=2, robust=True), sd, img.double(),
        labels, 0.1)
    m = m.to(DEV)
    set_mode(m, dtype)
    lg, ls, gr = model_loss_and_grads(m, img.to(DEV), labels.to(DEV), 0.1)
    if dtype == torch.float32:
        assert O.rel_l2(lg, ref_logits) < CHECK_REL
        worst, key = compare_grads(gr, ref_grads, O.rel_l2)
        assert worst < CHECK_REL, (key, worst)
    else:
        assert O.cosine(lg, ref_logits) > BF16_COS
        worst, key = compare_grads(gr, ref_grads, O.cosine)
        assert worst > BF16_COS, (key, worst)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("pool", ["cls", "mean"])
def test_readme_vit_matches_reference_golden(golden_dir, dtype, pool):
    """README `ViT` against golden vectors produced by the reference's runnable class of that structure
    (vit_with_patch_dropout.ViT with patch_dropout = 0; tests/golden/make_golden.py): logits, loss, every gradient.  The
    class-token row of pos_embedding has no counterpart in that class (zero there) and is left out of the comparison."""
    sd, grads, img, labels, logits, loss = load_readme_golden(os.path.join(golden_dir, "readme_vit.npz"), pool)
    m = V.ViT(**README_CFG, pool=pool)
    m.load_state_dict(sd)
    m = m.to(DEV)
    set_mode(m, dtype)
    lg, ls, gr = model_loss_and_grads(m, img.to(DEV), labels.to(DEV), 0.1)
    assert set(gr) == set(grads)
    gr["pos_embedding"], grads["pos_embedding"] = gr["pos_embedding"][:, 1:], grads["pos_embedding"][:, 1:]
    if dtype == torch.float32:
        assert O.rel_l2(lg, logits) < CHECK_REL and abs(ls - loss) < 1e-4
        worst, key = compare_grads(gr, grads, O.rel_l2)
        assert worst < CHECK_REL, (key, worst)
    else:
        assert O.cosine(lg, logits) > BF16_COS
        worst, key = compare_grads(gr, grads, O.cosine)
        assert worst > BF16_COS, (key, worst)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("pool", ["cls", "mean"])
def test_readme_vit_matches_restatement(dtype, pool):
    """README `ViT` API (README.md:67-111) at a second shape with a non-zero class-token positional row; the checker is
    the oracle's restatement, itself pinned against the reference class (tests/test_oracle.py, readme_vit.npz)."""
    kw = dict(image_size=64, patch_size=16, num_classes=40, dim=128, depth=2, heads=4, mlp_dim=256, dim_head=32)
    m = V.ViT(**kw, pool=pool, dropout=0.0, emb_dropout=0.0)
    randomize_(m, 33)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(9)
    img = torch.randn(3, 3, 64, 64, generator=g)
    labels = torch.randint(0, 40, (3,), generator=g)
    ref_logits, _, ref_grads = O.loss_and_grads(
        lambda s_, x: O.readme_vit_forward(s_, x, patch_size=16, heads=4, dim_head=32, pool=pool), sd, img.double(),
        labels, 0.1)
    m = m.to(DEV)
    set_mode(m, dtype)
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



def _library_mask(seed, layer, site, shape, p):
    """The keep/scale mask libnrvit draws for (seed, layer, site): nrv_dropout applied to ones."""
    import ctypes as C
    from vit_pytorch_robust import _abi
    n = 1
    for d in shape:
        n *= d
    n_pad = (n + 7) // 8 * 8
    ones = torch.ones(n_pad, device=DEV, dtype=torch.float32)
    out = torch.empty_like(ones)
    _abi.check(_abi.init(DEV).nrv_dropout(ones.data_ptr(), None, out.data_ptr(), n_pad, _abi.NRV_F32, p,
                                       C.c_ulonglong(seed), layer, site, _abi.stream_ptr()), "nrv_dropout")
    return out[:n].view(shape).cpu()


def test_dropout_mask_statistics_and_determinism():
    p = 0.3
    a = _library_mask(1234, 0, 1, (64, 4096), p)
    keep = (a != 0).float().mean().item()
    assert abs(keep - (1 - p)) < 5e-3
    assert torch.allclose(a[a != 0], torch.tensor(1.0 / (1 - p)))
    assert torch.equal(a, _library_mask(1234, 0, 1, (64, 4096), p))            # pure function of its key
    for other in (_library_mask(1235, 0, 1, (64, 4096), p), _library_mask(1234, 1, 1, (64, 4096), p),
                  _library_mask(1234, 0, 2, (64, 4096), p), _library_mask(1234, -1, 3, (64, 4096), p)):
        agree = ((other != 0) == (a != 0)).float().mean().item()               # independent masks agree on p^2+(1-p)^2
        assert abs(agree - (p * p + (1 - p) * (1 - p))) < 1e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("p_attn", [0.0, 0.2], ids=["dropout", "dropout+attention_dropout"])
def test_visiontransformer_dropout_matches_oracle_given_the_same_masks(dtype, p_attn):
    """train()-mode forward/backward with dropout = 0.25 (vit.py:45,47,125,174-175) and optionally
    attention_dropout = 0.2 (vit.py:105-110).  nn.Dropout's random stream cannot be reproduced, so the oracle is
    handed the masks the library drew (a pure function of the seed the module reports) and logits + every
    gradient must then agree to the usual tolerance."""
    p = 0.25
    m = V.VisionTransformer(**VIT_CFG, dropout=p, attention_dropout=p_attn)
    randomize_(m, 21)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(4)
    img = torch.randn(4, 3, 32, 32, generator=g)
    labels = torch.randint(0, 10, (4,), generator=g)
    m = m.to(DEV)
    set_mode(m, dtype)
    m.train()
    lg, ls, gr = model_loss_and_grads(m, img.to(DEV), labels.to(DEV), 0.1)
    req = m._nrv.last_dropout
    assert req is not None and req["p"] == p and req["p_emb"] == p and req["p_attn"] == p_attn

    def drop(t, layer, site):
        prob = p_attn if site == O.DROP_ATTN_PROB else p
        return t if prob == 0.0 else t * _library_mask(req["seed"], layer, site, tuple(t.shape), prob).to(t.dtype)

    ref_logits, _, ref_grads = O.loss_and_grads(
        lambda s_, x: O.vision_transformer_forward(s_, x, patch_size=8, num_heads=2, drop=drop), sd, img.double(), labels, 0.1)
    assert set(gr) == set(ref_grads)
    if dtype == torch.float32:
        assert O.rel_l2(lg, ref_logits) < CHECK_REL
        worst, key = compare_grads(gr, ref_grads, O.rel_l2)
        assert worst < CHECK_REL, (key, worst)
    else:
        assert O.cosine(lg, ref_logits) > BF16_COS
        worst, key = compare_grads(gr, ref_grads, O.cosine)
        assert worst > BF16_COS, (key, worst)
    # a second forward draws a new seed; eval() is the identity
    lg2, _, _ = model_loss_and_grads(m, img.to(DEV), labels.to(DEV), 0.1)
    assert m._nrv.last_dropout["seed"] != req["seed"] and not torch.equal(lg, lg2)
    m.eval()
    with torch.no_grad():
        e1, e2 = m(img.to(DEV)), m(img.to(DEV))
    assert m._nrv.last_dropout is None and torch.equal(e1, e2)


@pytest.mark.parametrize("patch,hidden,heads,batch", [(16, 128, 2, 3), (14, 160, 2, 2), (32, 96, 3, 5)],
                         ids=["197tok_dh64", "257tok_dh80", "50tok_dh32"])
def test_attention_dropout_runs_on_the_tensor_cores_and_matches_oracle(patch, hidden, heads, batch):
    """attention_dropout > 0 (vit.py:105-110; README ViT `dropout`) in bf16 stays on the tcgen05 path: the general
    attention kernels draw the mask themselves (forward: Philox call shared by four keys; backward: regenerated).
    attn_impl = TC makes the library raise instead of taking the CUDA-core kernels, so a green run proves which path
    ran.  The oracle receives the masks of the same (seed, layer, site) stream, drawn by nrv_dropout."""
    from vit_pytorch_robust import _abi
    p_attn = 0.3
    m = V.VisionTransformer(image_size=224, patch_size=patch, num_layers=2, num_heads=heads, hidden_dim=hidden,
                            mlp_dim=2 * hidden, num_classes=10, attention_dropout=p_attn)
    randomize_(m, 55)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(8)
    img = torch.randn(batch, 3, 224, 224, generator=g)
    labels = torch.randint(0, 10, (batch,), generator=g)
    m = m.to(DEV)
    set_mode(m, torch.bfloat16)
    m._nrv.attn_impl = _abi.ATTN_IMPL_TC
    m.train()
    lg, ls, gr = model_loss_and_grads(m, img.to(DEV), labels.to(DEV), 0.1)
    req = m._nrv.last_dropout
    assert req is not None and req["p_attn"] == p_attn and req["p"] == 0.0

    def drop(t, layer, site):
        if site != O.DROP_ATTN_PROB:
            return t
        return t * _library_mask(req["seed"], layer, site, tuple(t.shape), p_attn).to(t.dtype)

    ref_logits, _, ref_grads = O.loss_and_grads(
        lambda s_, x: O.vision_transformer_forward(s_, x, patch_size=patch, num_heads=heads, drop=drop), sd, img.double(), labels, 0.1)
    assert O.cosine(lg, ref_logits) > BF16_COS
    worst, key = compare_grads(gr, ref_grads, O.cosine)
    assert worst > BF16_COS, (key, worst)
    # the CUDA-core kernels draw the same masks: with the same seed both paths give the same step up to bf16 rounding
    torch.manual_seed(77)
    lg_tc, _, gr_tc = model_loss_and_grads(m, img.to(DEV), labels.to(DEV), 0.1)
    seed_tc = m._nrv.last_dropout["seed"]
    m._nrv.attn_impl = _abi.ATTN_IMPL_SIMT
    torch.manual_seed(77)
    lg_sm, _, gr_sm = model_loss_and_grads(m, img.to(DEV), labels.to(DEV), 0.1)
    assert m._nrv.last_dropout["seed"] == seed_tc
    assert O.cosine(lg_tc, lg_sm) > 0.9995
    worst, key = compare_grads(gr_tc, gr_sm, O.cosine)
    assert worst > 0.9995, (key, worst)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_readme_vit_dropout_matches_restatement(dtype):
    """README ViT(dropout=0.2, emb_dropout=0.1): `dropout` acts on the attention probabilities, after to_out and
    twice in the FeedForward; `emb_dropout` after the positional embedding (README.md:100-105).  The restatement (whose
    dropout sites are pinned against the reference class on the CPU, tests/test_oracle.py) receives the library's masks."""
    p, pe = 0.2, 0.1
    kw = dict(image_size=64, patch_size=16, num_classes=40, dim=128, depth=2, heads=4, mlp_dim=256, dim_head=32)
    m = V.ViT(**kw, dropout=p, emb_dropout=pe)
    randomize_(m, 34)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(10)
    img = torch.randn(3, 3, 64, 64, generator=g)
    labels = torch.randint(0, 40, (3,), generator=g)
    m = m.to(DEV)
    set_mode(m, dtype)
    m.train()
    lg, ls, gr = model_loss_and_grads(m, img.to(DEV), labels.to(DEV), 0.1)
    req = m._nrv.last_dropout

    def drop(t, layer, site):
        prob = pe if site == O.DROP_EMB else p
        return t * _library_mask(req["seed"], layer, site, tuple(t.shape), prob).to(t.dtype)

    ref_logits, _, ref_grads = O.loss_and_grads(
        lambda s_, x: O.readme_vit_forward(s_, x, patch_size=16, heads=4, dim_head=32, drop=drop), sd, img.double(), labels, 0.1)
    assert set(gr) == set(ref_grads)
    if dtype == torch.float32:
        assert O.rel_l2(lg, ref_logits) < CHECK_REL
        worst, key = compare_grads(gr, ref_grads, O.rel_l2)
        assert worst < CHECK_REL, (key, worst)
    else:
        assert O.cosine(lg, ref_logits) > BF16_COS
        worst, key = compare_grads(gr, ref_grads, O.cosine)
        assert worst > BF16_COS, (key, worst)


def test_readme_vit_shapes_and_dropout_contract():
    """README.md:67-86: v = ViT(..., dropout=0.1, emb_dropout=0.1); preds = v(img)  # (1, 1000), in eval() and train()."""
    v = V.ViT(image_size=256, patch_size=32, num_classes=1000, dim=1024, depth=6, heads=16, mlp_dim=2048,
              dropout=0.1, emb_dropout=0.1).to(DEV)
    img = torch.randn(1, 3, 256, 256, device=DEV)
    v.eval()
    with torch.no_grad():
        preds = v(img)
    assert preds.shape == (1, 1000)
    v.train()
    out = v(img)
    out.float().sum().backward()
    assert out.shape == (1, 1000) and v.pos_embedding.grad is not None and torch.isfinite(v.pos_embedding.grad).all()
    sv = V.SimpleViT(image_size=256, patch_size=32, num_classes=1000, dim=1024, depth=6, heads=16, mlp_dim=2048).to(DEV)
    assert sv(img).shape == (1, 1000)
    # torchvision-style class (vit.py:181-196)
    tv = V.vit.VisionTransformer(image_size=64, patch_size=16, num_layers=2, num_heads=2, hidden_dim=128, mlp_dim=256,
                                 dropout=0.1, attention_dropout=0.1, num_classes=10).to(DEV)
    x = torch.randn(2, 3, 64, 64, device=DEV)
    tv.eval()
    with torch.no_grad():
        assert tv(x).shape == (2, 10)
    tv.train()
    assert tv(x).shape == (2, 10)
    # not implemented (and no fallback): attention dropout together with Sinkhorn attention
    rb = V.vit.VisionTransformer(image_size=64, patch_size=16, num_layers=1, num_heads=2, hidden_dim=128, mlp_dim=256,
                                 attention_dropout=0.1, num_classes=10, robust=True).to(DEV)
    rb.train()
    with pytest.raises(NotImplementedError):
        rb(x)
    with pytest.raises(ValueError):
        V.vit.VisionTransformer(image_size=64, patch_size=16, num_layers=1, num_heads=2, hidden_dim=128, mlp_dim=256, dropout=1.5)


def test_vit_b16_full_size_properties():
    """BASELINE.json configs[2] at FULL size (ViT-B/16, 224x224, 256 images per GPU, bf16).  The CPU oracle does not
    finish at this size, so the checks are size-independent properties of the reference graph:
      (1) an image's logits do not depend on what else is in the batch (bit-exact: per-row arithmetic only);
      (2) the training-mode forward (activation stash) equals the inference forward (buffer rotation), bit-exact;
      (3) the gradient of the mean loss is linear in the batch: g(full) == g(first half) + g(second half), each half
          weighted 1/2 (only the fp32 summation order of the dW reductions differs);
      (4) bf16 logits agree with the fp32 check mode at the same size: cosine >= 0.999 (BASELINE.json tolerance)."""
    B = 256
    torch.manual_seed(0)
    m = V.vit_b_16()
    with torch.no_grad():
        m.heads.head.weight.normal_(std=0.02)
        m.class_token.normal_(std=0.02)
    m = m.to(DEV)
    g = torch.Generator().manual_seed(5)
    img = torch.randn(B, 3, 224, 224, generator=g).to(torch.bfloat16).to(DEV)
    labels = torch.randint(0, 1000, (B,), generator=g).to(DEV)
    m.eval()
    m._nrv.ln_fold_min_tokens = 0           # folded LayerNorm at every batch size (default: from 8192 tokens)
    with torch.no_grad():
        full = m(img)
        parts = torch.cat([m(img[i:i + 32]) for i in range(0, B, 32)])
    assert full.shape == (B, 1000) and torch.isfinite(full).all()
    assert torch.equal(full, parts)                                                    # (1)
    from vit_pytorch_robust import _abi
    assert m._nrv.ln_mode_infer == _abi.LN_FOLDED          # the inference default: LayerNorm folded into the QKV / FC1 GEMMs
    m._nrv.ln_mode_infer = _abi.LN_SEPARATE
    with torch.no_grad():
        full_sep = m(img)
    # same function, other rounding points: each is held to the BASELINE tolerance against the fp32 check mode in (4)
    assert O.cosine(full_sep, full) > 0.998 and not torch.equal(full_sep, full)
    m.train()
    m.zero_grad(set_to_none=True)
    assert m._nrv.ln_mode_train == _abi.LN_SEPARATE
    assert torch.equal(m(img).detach(), full_sep)                                      # (2) stand-alone LayerNorm kernels
    m._nrv.ln_mode_train = _abi.LN_FOLDED
    m.zero_grad(set_to_none=True)
    out = m(img)
    assert torch.equal(out.detach(), full)                                             # (2) folded LayerNorm
    V.softmax_cross_entropy(out, labels, 0.1).backward()
    g_full = m._nrv.flat_grad.clone()
    m._nrv.flat_grad.zero_()
    for lo in (0, B // 2):
        (0.5 * V.softmax_cross_entropy(m(img[lo:lo + B // 2]), labels[lo:lo + B // 2], 0.1)).backward()
    g_halves = m._nrv.flat_grad.clone()
    assert g_full.abs().max() > 0
    assert O.rel_l2(g_halves, g_full) < 1e-4                                           # (3)
    set_mode(m, torch.float32)
    m.eval()
    with torch.no_grad():
        ref = m(img[:64].float())
    assert O.cosine(full[:64], ref) > BF16_COS                                         # (4) folded LayerNorm
    assert O.rel_l2(full[:64], ref) < 3e-2
    assert O.cosine(full_sep[:64], ref) > BF16_COS                                     # (4) stand-alone LayerNorm kernels
    print("cosine vs fp32 check mode: folded %.6f  separate %.6f" % (O.cosine(full[:64], ref), O.cosine(full_sep[:64], ref)))


def test_small_batch_inference_replays_a_cuda_graph():
    """eval() + no_grad forwards of small batches are launch-bound: nrv_vit_forward is captured once per batch size and
    replayed.  The replay must follow new inputs and in-place parameter updates, bit-exactly equal to the direct call."""
    m = V.VisionTransformer(**VIT_CFG)
    randomize_(m, 8)
    m = m.to(DEV).eval()
    g = torch.Generator().manual_seed(12)
    xs = [torch.randn(4, 3, 32, 32, generator=g).to(DEV) for _ in range(3)]
    eng = m._nrv

    def direct(x):
        keep, eng.graph_max_tokens = eng.graph_max_tokens, 0
        try:
            with torch.no_grad():
                return m(x)
        finally:
            eng.graph_max_tokens = keep

    with torch.no_grad():
        a0 = m(xs[0])            # captures
        a1 = m(xs[1])            # replays with a new input
    assert len(eng._graphs) == 1
    assert torch.equal(a0, direct(xs[0])) and torch.equal(a1, direct(xs[1])) and not torch.equal(a0, a1)
    with torch.no_grad():        # in-place parameter update (what an optimiser step does): same graph, new weights
        for p in m.parameters():
            p.mul_(1.01)
        a2 = m(xs[2])
    assert len(eng._graphs) == 1
    assert torch.equal(a2, direct(xs[2]))
    with torch.no_grad():
        b = m(xs[0][:2])         # another batch size: its own graph
    assert len(eng._graphs) == 2 and torch.equal(b, direct(xs[0][:2]))


def test_vit_h14_shape_trains_through_the_general_tcgen05_attention_backward():
    """257 tokens with dh = 80 (ViT-H/14): outside the fused training kernels, the general tcgen05 forward and backward
    (attention_fwd_big.cu / attention_bwd_big.cu) carry it -- logits and gradients match the oracle."""
    kw = dict(image_size=224, patch_size=14, num_layers=1, num_heads=2, hidden_dim=160, mlp_dim=320, num_classes=12)
    m = V.VisionTransformer(**kw)
    randomize_(m, 77)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(2)
    img = torch.randn(2, 3, 224, 224, generator=g)
    labels = torch.randint(0, 12, (2,), generator=g)
    ref_logits, _, ref_grads = O.loss_and_grads(
        lambda s_, x: O.vision_transformer_forward(s_, x, patch_size=14, num_heads=2), sd, img.double(), labels, 0.1)
    m = m.to(DEV)
    set_mode(m, torch.bfloat16)
    lg, ls, gr = model_loss_and_grads(m, img.to(DEV), labels.to(DEV), 0.1)
    assert O.cosine(lg, ref_logits) > BF16_COS
    worst, key = compare_grads(gr, ref_grads, O.cosine)
    assert worst > BF16_COS, (key, worst)
    # fp32 check mode at this shape: the CUDA-core attention backward keeps two of the four head matrices resident at a time
    # (attn_bwd_simt_stream_kernel); rel <= 1e-3 is the bound north_star demands, the measured error is ~1e-6
    set_mode(m, torch.float32)
    m.zero_grad(set_to_none=True)
    lg, ls, gr = model_loss_and_grads(m, img.to(DEV), labels.to(DEV), 0.1)
    assert O.rel_l2(lg, ref_logits) < CHECK_REL
    worst, key = compare_grads(gr, ref_grads, O.rel_l2)
    assert worst < CHECK_REL, (key, worst)



# ---------------------------------------------------------------- introspection (recorder.py / extractor.py)
def _ref_wrapper_source(name):
    """The reference's Recorder / Extractor classes, when /root/reference is present (this container); the GPU box
    uses the equivalent hook registrations below."""
    import os
    path = os.path.join("/root/reference/vit_pytorch_robust", name)
    return path if os.path.exists(path) else None


@pytest.mark.parametrize("robust", [False, True])
def test_recorder_style_hooks_on_attend_see_the_attention_probabilities(robust):
    """recorder.py:28-31 registers a forward hook on every Attention.attend and stacks the outputs to [B, L, H, N, N].
    The fused attention never materialises them; with a hook present the model recomputes them from the stashed projections
    (nrv_attn_probs) and calls `attend` so the hook fires.  Checked against the oracle's probabilities."""
    torch.manual_seed(0)
    cfg = dict(image_size=32, patch_size=8, num_classes=10, dim=64, depth=3, heads=2, mlp_dim=128, robust=robust)
    m = V.SimpleViT(**cfg)
    randomize_(m, 3)
    sd = {k: v.detach().clone().double() for k, v in m.state_dict().items()}
    m = m.to(DEV).eval()
    x = torch.randn(4, 3, 32, 32)
    recordings = []
    hooks = [attn.attend.register_forward_hook(lambda mod, inp, out: recordings.append(out.clone().detach()))
             for attn, _ in m.transformer.layers]
    with torch.no_grad():
        logits = m(x.to(DEV))
    attns = torch.stack(recordings, dim=1)
    assert attns.shape == (4, 3, 2, 16, 16)
    ref_logits, ref_attn = O.simple_vit_forward(sd, x.double(), patch_size=8, heads=2, dim_head=64, robust=robust,
                                                return_attn=True)
    assert O.cosine(logits, ref_logits) > 0.999
    assert O.rel_l2(attns.cpu(), torch.stack(ref_attn, dim=1)) < 2e-2          # bf16 projections feed the probabilities
    assert torch.allclose(attns.sum(-1), torch.ones_like(attns.sum(-1)), atol=1e-4)
    for h in hooks:
        h.remove()
    recordings.clear()
    with torch.no_grad():
        again = m(x.to(DEV))
    assert not recordings and torch.equal(again, logits)                       # no hooks: the plain fused path, same result


def test_extractor_style_hook_on_transformer_returns_the_tokens():
    """extractor.py:50-59 hooks `vit.transformer` and returns its output tokens [B, N, D] next to the prediction."""
    torch.manual_seed(0)
    m = V.SimpleViT(image_size=32, patch_size=8, num_classes=10, dim=64, depth=2, heads=2, mlp_dim=128)
    randomize_(m, 4)
    sd = {k: v.detach().clone().double() for k, v in m.state_dict().items()}
    m = m.to(DEV)
    x = torch.randn(3, 3, 32, 32)
    got = {}
    h = m.transformer.register_forward_hook(lambda mod, inp, out: got.update(inp=inp[0], out=out))
    logits = m(x.to(DEV))          # grad mode: the stash is the training stash
    logits.float().sum().backward()
    h.remove()
    ref_logits = O.simple_vit_forward(sd, x.double(), patch_size=8, heads=2, dim_head=64)
    ref_tokens = O.simple_vit_forward(sd, x.double(), patch_size=8, heads=2, dim_head=64, return_tokens=True)
    assert got["out"].shape == (3, 16, 64) and got["inp"].shape == (3, 16, 64)
    assert O.cosine(got["out"], ref_tokens) > 0.999
    assert O.cosine(logits, ref_logits) > 0.999
    assert all(p.grad is not None for p in m.parameters())


def test_hooks_on_fused_parameter_holders_raise():
    m = V.SimpleViT(image_size=32, patch_size=8, num_classes=10, dim=64, depth=2, heads=2, mlp_dim=128).to(DEV)
    h = m.transformer.layers[0][1].register_forward_hook(lambda *a: None)
    with pytest.raises(NotImplementedError, match="forward hook"):
        m(torch.randn(1, 3, 32, 32, device=DEV))
    h.remove()
    m(torch.randn(1, 3, 32, 32, device=DEV))


def test_vision_transformer_block_hooks_see_the_residual_stream():
    torch.manual_seed(0)
    m = V.VisionTransformer(**VIT_CFG)
    randomize_(m, 5)
    sd = {k: v.detach().clone().double() for k, v in m.state_dict().items()}
    m = m.to(DEV).eval()
    x = torch.randn(2, 3, VIT_CFG["image_size"], VIT_CFG["image_size"])
    seen = {}
    blk = m.encoder.layers[1]
    h1 = blk.register_forward_hook(lambda mod, inp, out: seen.update(blk_in=inp[0], blk_out=out))
    h2 = m.encoder.layers.register_forward_hook(lambda mod, inp, out: seen.update(enc_out=out))
    with torch.no_grad():
        logits = m(x.to(DEV))
    h1.remove(); h2.remove()
    ref_logits, streams = O.vision_transformer_forward(sd, x.double(), patch_size=VIT_CFG["patch_size"],
                                                       num_heads=VIT_CFG["num_heads"], return_streams=True)
    assert O.cosine(logits, ref_logits) > 0.999
    assert O.cosine(seen["blk_in"], streams[1]) > 0.999 and O.cosine(seen["blk_out"], streams[2]) > 0.999
    assert O.cosine(seen["enc_out"], streams[-1]) > 0.999


# ---------------------------------------------------------------- checkpoint interchange
def test_evaluation_style_checkpoint_roundtrip(tmp_path):
    """examples/evaluation.py:129-140: heads.head = Identity, torch.load(ckpt)["model"], strip `module.`, load_state_dict,
    requires_grad_(False), then linear probes on the frozen features.  The checkpoint here is written from a torchvision
    VisionTransformer wrapped the way torch DDP names its keys; the frozen fused encoder must reproduce torchvision's features."""
    from torchvision.models.vision_transformer import VisionTransformer as TV
    torch.manual_seed(0)
    tv = TV(image_size=32, patch_size=8, num_layers=2, num_heads=2, hidden_dim=64, mlp_dim=128, num_classes=10)
    with torch.no_grad():
        for p in tv.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    tv.heads.head = torch.nn.Identity()
    path = tmp_path / "final.ckpt"
    torch.save({"model": {"module." + k: v for k, v in tv.state_dict().items()}, "epoch": 3}, path)

    model = V.VisionTransformer(**VIT_CFG)
    model.heads.head = torch.nn.Identity()                                   # evaluation.py:129-131
    ckpt = torch.load(path, map_location="cpu")["model"]
    ckpt = {k.replace("module.", ""): v for k, v in ckpt.items()}
    model.load_state_dict(ckpt)                                              # strict: keys and shapes are the reference's
    model.requires_grad_(False)
    model = model.to(DEV).eval()
    x = torch.randn(5, 3, 32, 32)
    with torch.no_grad():
        want = tv.eval()(x)
        got = model(x.to(DEV))
    assert got.shape == want.shape == (5, 64)
    assert O.cosine(got, want) > 0.999
    # a probe trains on top of the frozen encoder; the encoder keeps no stash and receives no gradient
    probe = torch.nn.Linear(64, 10).to(DEV)
    loss = torch.nn.functional.cross_entropy(probe(model(x.to(DEV)).float()), torch.randint(0, 10, (5,), device=DEV))
    loss.backward()
    assert probe.weight.grad is not None and all(p.grad is None for p in model.parameters())


def test_legacy_torchvision_mlp_keys_are_remapped():
    """vit.py:55-84 (MLPBlock._load_from_state_dict): checkpoints written before torchvision 0.13 name the MLP Linears
    linear_1 / linear_2; they load into mlp.0 / mlp.3 when the metadata carries no version."""
    m = V.VisionTransformer(**VIT_CFG)
    sd = m.state_dict()
    legacy = type(sd)()
    for k, v in sd.items():
        legacy[k.replace(".mlp.0.", ".mlp.linear_1.").replace(".mlp.3.", ".mlp.linear_2.")] = v.clone() + 1.0
    assert any("linear_1" in k for k in legacy)
    m2 = V.VisionTransformer(**VIT_CFG)
    missing, unexpected = m2.load_state_dict(legacy, strict=True)
    assert not missing and not unexpected
    for k, v in m2.state_dict().items():
        assert torch.equal(v, sd[k] + 1.0), k


def test_load_state_dict_after_first_use_refreshes_the_tensor_core_copy():
    """Parameters are views of one flat fp32 buffer with a bf16 shadow for the tensor cores: loading new weights into a model
    that has already run must be visible in the next forward (evaluation sweeps load checkpoint after checkpoint)."""
    torch.manual_seed(1)
    m = V.VisionTransformer(**VIT_CFG)
    randomize_(m, 6)
    m = m.to(DEV).eval()
    x = torch.randn(3, 3, 32, 32, device=DEV)
    with torch.no_grad():
        a = m(x)
    other = V.VisionTransformer(**VIT_CFG)
    randomize_(other, 7)
    m.load_state_dict(other.state_dict())
    sd = {k: v.detach().clone().double().cpu() for k, v in other.state_dict().items()}
    with torch.no_grad():
        b = m(x)
    want = O.vision_transformer_forward(sd, x.double().cpu(), patch_size=VIT_CFG["patch_size"], num_heads=VIT_CFG["num_heads"])
    assert O.cosine(b, want) > 0.999 and O.cosine(a, want) < 0.99
