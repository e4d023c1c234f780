"""Generates the golden fixtures in this directory by RUNNING THE REFERENCE (in the build container,
where /root/reference exists).  Re-run with:  python tests/golden/make_golden.py

Fixtures (small, float32, seeded):
  simplevit_{softmax,robust}.npz — reference vit_pytorch_robust.simple_vit.SimpleViT (as shipped):
        state_dict, input, labels, logits, loss, gradient of every parameter (CE, label smoothing 0.1)
  visiontransformer_softmax.npz  — torchvision twin of reference vit.VisionTransformer loaded with
        the REFERENCE constructor's state_dict (head / class token re-randomised: the reference
        zero-initialises them, vit.py:247,304-306, which would make every gradient zero)
  posemb_sincos.npz              — reference posemb_sincos_2d on an 8x8 grid, dim 512
  sinkhorn.npz                   — reference utils.SinkhornAttention on seeded 14x14 and peaky 64x64 inputs
  readme_vit.npz                 — reference vit_with_patch_dropout.ViT(patch_dropout=0) (the runnable in-tree class with
        the README `ViT` structure), cls and mean pooling: state_dict under the REFERENCE's keys, logits, loss, gradients
        (python tests/golden/make_golden.py readme_vit  regenerates this one only)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import ref_loader  # noqa: E402

SIMPLE_CFG = dict(image_size=32, patch_size=8, num_classes=10, dim=64, depth=2, heads=2, mlp_dim=128, dim_head=32)
VIT_CFG = dict(image_size=32, patch_size=8, num_layers=2, num_heads=2, hidden_dim=64, mlp_dim=128, num_classes=10)


def to_np(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}


def run(model, img, labels, ls):
    logits = model(img)
    loss = torch.nn.functional.cross_entropy(logits, labels, label_smoothing=ls)
    model.zero_grad()
    loss.backward()
    grads = {"grad::" + k: p.grad.detach().cpu().numpy() for k, p in model.named_parameters() if p.grad is not None}
    return logits.detach().cpu().numpy(), loss.item(), grads


README_CFG = dict(image_size=32, patch_size=8, num_classes=10, dim=64, depth=2, heads=2, mlp_dim=128, dim_head=32)


def make_readme_vit(ref):
    out = {}
    torch.manual_seed(1234)
    img = torch.randn(4, 3, 32, 32)
    labels = torch.randint(0, 10, (4,))
    for pool in ("cls", "mean"):
        torch.manual_seed(21)
        m = ref.vit_with_patch_dropout.ViT(**README_CFG, pool=pool, patch_dropout=0.)
        logits, loss, grads = run(m, img, labels, 0.1)
        out.update({pool + "::param::" + k: v for k, v in to_np(m.state_dict()).items()})
        out.update({pool + "::" + k: v for k, v in grads.items()})
        out.update({pool + "::logits": logits, pool + "::loss": np.float32(loss)})
    out.update(img=img.numpy(), labels=labels.numpy())
    np.savez_compressed(os.path.join(HERE, "readme_vit.npz"), **out)


def main():
    ref = ref_loader.load_reference()
    assert ref is not None, "reference tree not found"
    if len(sys.argv) > 1 and sys.argv[1] == "readme_vit":
        make_readme_vit(ref)
        return
    make_readme_vit(ref)
    torch.manual_seed(1234)
    img = torch.randn(4, 3, 32, 32)
    labels = torch.randint(0, 10, (4,))

    for robust in (False, True):
        torch.manual_seed(7)
        m = ref.simple_vit.SimpleViT(**SIMPLE_CFG, robust=robust)
        logits, loss, grads = run(m, img, labels, 0.1)
        out = {"param::" + k: v for k, v in to_np(m.state_dict()).items()}
        out.update(grads)
        out.update(img=img.numpy(), labels=labels.numpy(), logits=logits, loss=np.float32(loss))
        np.savez_compressed(os.path.join(HERE, "simplevit_%s.npz" % ("robust" if robust else "softmax")), **out)

    torch.manual_seed(11)
    rv = ref.vit.VisionTransformer(**VIT_CFG)          # reference constructor (forward is broken as shipped)
    sd = rv.state_dict()
    g = torch.Generator().manual_seed(3)
    sd["heads.head.weight"] = torch.randn(sd["heads.head.weight"].shape, generator=g) * 0.1
    sd["heads.head.bias"] = torch.randn(sd["heads.head.bias"].shape, generator=g) * 0.1
    sd["class_token"] = torch.randn(sd["class_token"].shape, generator=g) * 0.1
    twin = ref_loader.torchvision_twin(**VIT_CFG)
    missing = twin.load_state_dict(sd)
    assert not missing.missing_keys and not missing.unexpected_keys
    logits, loss, grads = run(twin, img, labels, 0.1)
    out = {"param::" + k: v for k, v in to_np(twin.state_dict()).items()}
    out.update(grads)
    out.update(img=img.numpy(), labels=labels.numpy(), logits=logits, loss=np.float32(loss))
    np.savez_compressed(os.path.join(HERE, "visiontransformer_softmax.npz"), **out)

    pe = ref.simple_vit.posemb_sincos_2d(torch.zeros(1, 8, 8, 512))
    np.savez_compressed(os.path.join(HERE, "posemb_sincos.npz"), pe=pe.numpy())

    torch.manual_seed(5)
    a = torch.rand(14, 14)
    b = 3 * torch.randn(2, 3, 64, 64)
    sk = ref.utils.SinkhornAttention(-1)
    np.savez_compressed(os.path.join(HERE, "sinkhorn.npz"), a=a.numpy(), a_out=sk(a).numpy(), b=b.numpy(),
                        b_out=sk(b).numpy())
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
