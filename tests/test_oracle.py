"""Pins the oracle (oracle/vit_oracle.py) against golden vectors generated from the reference
(tests/golden/make_golden.py) and, when the reference tree is present, against the reference live."""
import os

import numpy as np
import pytest
import torch

import ref_loader
import vit_oracle as O
from helpers import SIMPLE_CFG, VIT_CFG, load_golden


@pytest.mark.parametrize("name,robust", [("softmax", False), ("robust", True)])
def test_simplevit_oracle_matches_golden(golden_dir, name, robust):
    sd, grads, img, labels, logits, loss = load_golden(os.path.join(golden_dir, "simplevit_%s.npz" % name))
    lg, ls, gr = O.loss_and_grads(O.simple_vit_forward, sd, img, labels, 0.1, patch_size=8, heads=2, dim_head=32,
                                  robust=robust)
    assert O.rel_l2(lg, logits) < 2e-6
    assert abs(ls.item() - loss) < 2e-6
    assert set(gr) == set(grads)
    for k in grads:
        assert O.rel_l2(gr[k], grads[k]) < 5e-6, k


def test_visiontransformer_oracle_matches_golden(golden_dir):
    sd, grads, img, labels, logits, loss = load_golden(os.path.join(golden_dir, "visiontransformer_softmax.npz"))
    lg, ls, gr = O.loss_and_grads(O.vision_transformer_forward, sd, img, labels, 0.1, patch_size=8, num_heads=2)
    assert O.rel_l2(lg, logits) < 5e-6
    assert abs(ls.item() - loss) < 2e-6
    assert set(gr) == set(grads)
    for k in grads:
        assert O.rel_l2(gr[k], grads[k]) < 1e-5, k


def test_posemb_golden(golden_dir):
    pe = np.load(os.path.join(golden_dir, "posemb_sincos.npz"))["pe"]
    got = O.posemb_sincos_2d(8, 8, 512).numpy()
    assert np.abs(got - pe).max() == 0.0
    # known-answer facts recorded in SURVEY.md section 8c
    np.testing.assert_allclose(pe[9, 0:3], [0.8415, 0.8016, 0.7611], atol=1e-4)
    np.testing.assert_allclose(pe[9, 128:131], [0.5403, 0.5978, 0.6487], atol=1e-4)
    assert np.array_equal(pe[9, 0:256], pe[9, 256:512])  # token 9 = (y=1, x=1)


def test_patch_index_maps():
    x = torch.arange(2 * 3 * 8 * 12, dtype=torch.float32).reshape(2, 3, 8, 12)
    P = 4
    y = O.patchify_p1p2c(x, P, P).reshape(2, 2, 3, P * P * 3)
    for (h, w, p1, p2, c) in [(0, 0, 0, 0, 0), (1, 2, 3, 1, 2), (0, 1, 2, 3, 1)]:
        assert y[0, h, w, (p1 * P + p2) * 3 + c] == x[0, c, h * P + p1, w * P + p2]
    z = O.patchify_cp1p2(x, P, P).reshape(2, 2, 3, 3 * P * P)
    for (h, w, p1, p2, c) in [(0, 0, 0, 0, 0), (1, 2, 3, 1, 2), (0, 1, 2, 3, 1)]:
        assert z[1, h, w, (c * P + p1) * P + p2] == x[1, c, h * P + p1, w * P + p2]
    # conv weight <-> linear weight equivalence
    conv = torch.nn.Conv2d(3, 5, P, P)
    ref = conv(x).flatten(2).transpose(1, 2)
    got = O.patchify_cp1p2(x, P, P) @ conv.weight.reshape(5, -1).t() + conv.bias
    assert torch.allclose(ref, got, atol=1e-3, rtol=1e-5)


def test_sinkhorn_golden_and_contract(golden_dir):
    s = np.load(os.path.join(golden_dir, "sinkhorn.npz"))
    a = O.sinkhorn3(torch.from_numpy(s["a"]).softmax(-1))
    assert np.abs(a.numpy() - s["a_out"]).max() < 1e-7
    assert torch.allclose(a.sum(-1), torch.ones(14), atol=1e-6)
    assert torch.allclose(a.sum(-2), torch.ones(14), atol=1e-5)
    b = O.sinkhorn3(torch.from_numpy(s["b"]).softmax(-1))
    assert np.abs(b.numpy() - s["b_out"]).max() < 1e-7
    # exactly three iterations is the contract, not convergence: peaky logits leave column sums off
    assert (b.sum(-2) - 1).abs().max() > 1e-2


def test_cross_entropy_and_adamw_restatements():
    torch.manual_seed(0)
    z = torch.randn(6, 11)
    y = torch.randint(0, 11, (6,))
    for ls in (0.0, 0.1, 0.8):
        assert torch.allclose(O.cross_entropy(z, y, ls), torch.nn.functional.cross_entropy(z, y, label_smoothing=ls),
                              atol=1e-6)
    p = torch.randn(50, dtype=torch.float64)
    tp = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([tp], lr=2e-3, weight_decay=0.05)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step in range(1, 4):
        g = torch.randn(50, dtype=torch.float64)
        tp.grad = g.clone()
        opt.step()
        p, m, v = O.adamw_step(p, g, m, v, step, 2e-3, weight_decay=0.05)
        assert torch.allclose(p, tp.detach(), atol=1e-12)


@pytest.mark.skipif(ref_loader.find_reference_dir() is None, reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("robust", [False, True])
def test_simplevit_oracle_matches_live_reference(robust):
    ref = ref_loader.load_reference()
    torch.manual_seed(3)
    m = ref.simple_vit.SimpleViT(image_size=(24, 40), patch_size=(8, 4), num_classes=7, dim=48, depth=3, heads=3,
                                 mlp_dim=96, dim_head=16, robust=robust).double()
    img = torch.randn(3, 3, 24, 40, dtype=torch.float64)
    want = m(img)
    got = O.simple_vit_forward(m.state_dict(), img, patch_size=(8, 4), heads=3, dim_head=16, robust=robust)
    assert O.rel_l2(got, want) < 1e-12


@pytest.mark.skipif(ref_loader.find_reference_dir() is None, reason="reference tree not present (GPU box)")
def test_vit_state_dict_matches_reference_constructor():
    ref = ref_loader.load_reference()
    if ref.vit is None:
        pytest.skip("reference vit.py not importable with this torchvision")
    import vit_pytorch_robust as ours
    a = ref.vit.VisionTransformer(**VIT_CFG).state_dict()
    b = ours.VisionTransformer(**VIT_CFG).state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape for k in a)
    a = ref.simple_vit.SimpleViT(**SIMPLE_CFG).state_dict()
    b = ours.SimpleViT(**SIMPLE_CFG).state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape for k in a)


def _mask(key, shape, p):
    """Deterministic keep/scale mask for a dropout site, shared by the twin and the oracle."""
    g = torch.Generator().manual_seed(1000 + 17 * (key[0] + 1) + key[1])
    return (torch.rand(shape, generator=g, dtype=torch.float64) >= p).to(torch.float64) / (1.0 - p)


class _MaskDrop(torch.nn.Module):
    def __init__(self, key, p):
        super().__init__()
        self.key, self.p = key, p

    def forward(self, x):
        return x * _mask(self.key, tuple(x.shape), self.p).to(x.dtype) if self.training else x


def test_oracle_dropout_sites_match_the_torchvision_twin():
    """Where the oracle applies dropout (vit.py:45,47 MLP; :125 after attention; :174-175 embedding) is pinned against the
    class vit.py was copied from: every nn.Dropout of a train()-mode torchvision VisionTransformer is replaced by a
    module that multiplies by a seeded mask, and the oracle is handed the same masks."""
    from torchvision.models.vision_transformer import VisionTransformer as TV
    p = 0.3
    torch.manual_seed(3)
    tv = TV(**VIT_CFG, dropout=p).double()
    with torch.no_grad():
        for prm in tv.parameters():
            if float(prm.abs().sum()) == 0.0:
                prm.normal_(std=0.05)
    tv.encoder.dropout = _MaskDrop((-1, O.DROP_EMB), p)
    for i, blk in enumerate(tv.encoder.layers):
        blk.dropout = _MaskDrop((i, O.DROP_ATTN_OUT), p)
        blk.mlp[2] = _MaskDrop((i, O.DROP_FC1), p)
        blk.mlp[4] = _MaskDrop((i, O.DROP_FC2), p)
    tv.train()
    g = torch.Generator().manual_seed(5)
    img = torch.randn(3, 3, 32, 32, generator=g, dtype=torch.float64)
    labels = torch.randint(0, 10, (3,), generator=g)
    logits = tv(img)
    loss = torch.nn.functional.cross_entropy(logits, labels, label_smoothing=0.1)
    loss.backward()
    sd = {k: v.detach().clone() for k, v in tv.state_dict().items()}
    drop = lambda t, layer, site: t if site == O.DROP_ATTN_PROB else t * _mask((layer, site), tuple(t.shape), p)  # noqa: E731
    lg, ls, gr = O.loss_and_grads(lambda s_, x: O.vision_transformer_forward(s_, x, patch_size=8, num_heads=2, drop=drop),
                                  sd, img, labels, 0.1)
    assert O.rel_l2(lg, logits) < 1e-12 and abs(ls.item() - loss.item()) < 1e-12
    for k, prm in tv.named_parameters():
        assert O.rel_l2(gr[k], prm.grad) < 1e-10, k
    # the masks matter: without them the logits differ
    lg0, _, _ = O.loss_and_grads(lambda s_, x: O.vision_transformer_forward(s_, x, patch_size=8, num_heads=2), sd, img, labels, 0.1)
    assert O.rel_l2(lg0, logits) > 1e-3


def test_dropout_request_host_logic():
    from vit_pytorch_robust import engine as E
    assert E.dropout_request(False, p=0.5, p_emb=0.5, p_attn=0.5) is None           # eval(): identity
    assert E.dropout_request(True) is None                                          # p = 0: identity
    torch.manual_seed(11)
    a = E.dropout_request(True, p=0.1, p_emb=0.2, p_attn=0.3)
    torch.manual_seed(11)
    b = E.dropout_request(True, p=0.1, p_emb=0.2, p_attn=0.3)
    c = E.dropout_request(True, p=0.1, p_emb=0.2, p_attn=0.3)
    assert a == b and a["seed"] != c["seed"] and (a["p"], a["p_emb"], a["p_attn"]) == (0.1, 0.2, 0.3)
    with pytest.raises(NotImplementedError):
        E.dropout_request(True, p_attn=0.1, robust=True)                             # no fused Sinkhorn + dropout kernel
