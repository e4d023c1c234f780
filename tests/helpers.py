"""Shared helpers for the parity tests (oracle = checker, never the thing under test)."""
import numpy as np
import torch

import vit_oracle as O

SIMPLE_CFG = dict(image_size=32, patch_size=8, num_classes=10, dim=64, depth=2, heads=2, mlp_dim=128, dim_head=32)
VIT_CFG = dict(image_size=32, patch_size=8, num_layers=2, num_heads=2, hidden_dim=64, mlp_dim=128, num_classes=10)


def load_golden(path):
    z = np.load(path)
    sd = {k[7:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param::")}
    grads = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad::")}
    return sd, grads, torch.from_numpy(z["img"]), torch.from_numpy(z["labels"]), torch.from_numpy(z["logits"]), float(z["loss"])


README_CFG = dict(image_size=32, patch_size=8, num_classes=10, dim=64, depth=2, heads=2, mlp_dim=128, dim_head=32)


def load_readme_golden(path, pool):
    """readme_vit.npz: the reference's vit_with_patch_dropout.ViT(patch_dropout=0) under its own keys, translated to the
    README `ViT` keys (oracle key map).  Gradients are translated the same way; the class-token row of pos_embedding has
    no counterpart in the reference class (it stays zero there) and is dropped from the comparison by the callers."""
    z = np.load(path)
    pre = pool + "::"
    sd = O.readme_state_from_patch_dropout_vit(
        {k[len(pre) + 7:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre + "param::")})
    grads = O.readme_state_from_patch_dropout_vit(
        {k[len(pre) + 6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre + "grad::")})
    return (sd, grads, torch.from_numpy(z["img"]), torch.from_numpy(z["labels"]), torch.from_numpy(z[pre + "logits"]),
            float(z[pre + "loss"]))


def randomize_(model, seed=0, scale=1.0):
    """Seeded re-initialisation that leaves no parameter at zero (the reference zero-initialises
    heads.head and class_token, vit.py:247,304-306, which would make gradients vanish)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * (scale / fan_in ** 0.5))
            elif "norm" in name or "ln" in name or name.endswith("0.weight"):
                if name.endswith("weight"):
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
                else:
                    p.copy_(0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))


def model_loss_and_grads(model, img, labels, ls):
    """Runs the product model (CUDA) and returns logits, loss and parameter gradients on the CPU."""
    model.zero_grad(set_to_none=True)
    logits = model(img)
    loss = torch.nn.functional.cross_entropy(logits.float(), labels, label_smoothing=ls)
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().float().cpu().clone() for k, p in model.named_parameters() if p.grad is not None}
    return logits.detach().float().cpu(), loss.item(), grads


def compare_grads(got, ref, metric):
    worst = (None, None)
    for k, r in ref.items():
        assert k in got, "missing gradient for %s" % k
        v = metric(got[k], r)
        if worst[0] is None or (v > worst[0] if metric is O.rel_l2 else v < worst[0]):
            worst = (v, k)
    return worst
