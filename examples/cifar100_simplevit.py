"""CIFAR-100-shaped SimpleViT training: the SimpleViT branch of the reference's examples/CIFAR100.py (:71-80, commented out
as shipped) with that script's objective -- CutMix with probability cutmix_prob (:116-125), CE with label smoothing 0.1
(:127-136), AdamW (:90-97), linear warm-up + cosine schedule stepped per iteration (:99-112,160), grad_max_norm 5.0 (:192).
Shape = BASELINE.json configs[1]: 32x32 images, patch 4, dim 512, depth 6, heads 8.  Synthetic data.

  python examples/cifar100_simplevit.py --batch-size 1024 --steps-per-epoch 20 [--robust] [--fused-optimizer]
"""
import argparse

import numpy as np
import torch
from torch.optim.lr_scheduler import CosineAnnealingLR, LinearLR, SequentialLR

import omega_min as omega
from omega_min import V


def rand_bbox(size, lam):
    """vit_pytorch_robust/utils.py rand_bbox as CIFAR100.py:121 calls it: a box covering (1 - lam) of the image."""
    W, H = size[2], size[3]
    cut_rat = np.sqrt(1.0 - lam)
    cut_w, cut_h = int(W * cut_rat), int(H * cut_rat)
    cx, cy = np.random.randint(W), np.random.randint(H)
    return (np.clip(cx - cut_w // 2, 0, W), np.clip(cy - cut_h // 2, 0, H),
            np.clip(cx + cut_w // 2, 0, W), np.clip(cy + cut_h // 2, 0, H))


class Model(omega.Trainer):
    def initialize_train_loader(self):
        self.num_classes, self.image_size = omega.NAME_TO_CLASS[self.args.dataset]
        per_device = self.args.batch_size // self.args.world_size                     # CIFAR100.py:22
        return omega.synthetic_loader(self.args.steps_per_epoch, per_device, self.image_size, self.num_classes,
                                      self.this_device, seed=self.rank)

    def initialize_val_loader(self):
        return omega.synthetic_loader(2, self.args.batch_size // self.args.world_size, self.image_size, self.num_classes,
                                      self.this_device, seed=1000 + self.rank)

    def initialize_modules(self):
        a = self.args
        self.model = V.SimpleViT(image_size=self.image_size, patch_size=a.ps, num_classes=self.num_classes, dim=a.dim,
                                 depth=a.depth, heads=a.heads, mlp_dim=a.mlp_dim, robust=a.robust)   # CIFAR100.py:71-80

    def initialize_scheduler(self):                                                   # CIFAR100.py:99-112
        train_steps = len(self.train_loader)
        T1 = max(1, int(self.args.epochs * 0.1) * train_steps)
        T2 = max(1, (self.args.epochs - int(self.args.epochs * 0.1)) * train_steps)
        return SequentialLR(self.optimizer,
                            [LinearLR(self.optimizer, 1e-3, 1, total_iters=T1),
                             CosineAnnealingLR(self.optimizer, T_max=T2, eta_min=self.args.learning_rate * 0.05)],
                            milestones=[T1])

    def compute_loss(self):                                                           # CIFAR100.py:114-138
        x, y = self.data
        r = np.random.rand(1)
        if r < self.args.cutmix_prob:
            x = x.clone()
            rand_index = torch.randperm(x.size(0), device=x.device)
            lam = np.random.beta(self.args.beta, self.args.beta)
            bbx1, bby1, bbx2, bby2 = rand_bbox(x.size(), lam)
            x[:, :, bbx1:bbx2, bby1:bby2] = x[rand_index, :, bbx1:bbx2, bby1:bby2]
            lam = 1 - (bbx2 - bbx1) * (bby2 - bby1) / (x.size(-1) * x.size(-2))
        preds = self.model(x)
        if r < self.args.cutmix_prob:
            return (V.softmax_cross_entropy(preds, y, 0.1) * lam +
                    V.softmax_cross_entropy(preds, y[rand_index], 0.1) * (1 - lam))
        return V.softmax_cross_entropy(preds, y, 0.1)

    def before_eval_epoch(self):
        super().before_eval_epoch()
        self.accu, self.counter = 0.0, 0

    def eval_step(self):                                                              # CIFAR100.py:145-153
        x, y = self.data
        accu = self.model(x).argmax(1).eq(y).float().mean()
        if self.args.world_size > 1:
            torch.distributed.reduce(accu, dst=0)
        self.accu += accu.item()
        self.counter += 1

    def after_eval_epoch(self):
        super().after_eval_epoch()
        self.log_txt("eval_accuracies", accus=(self.accu / self.counter) / self.args.world_size)

    def after_train_step(self):
        self.scheduler.step()


def main(argv=None):
    parser = argparse.ArgumentParser(description="SimpleViT on CIFAR-100-shaped synthetic data")
    parser.add_argument("--beta", default=1.0, type=float)
    parser.add_argument("--cutmix_prob", default=0.0, type=float)
    parser.add_argument("--robust", action="store_true")
    for name, default in (("ps", 4), ("dim", 512), ("depth", 6), ("heads", 8), ("mlp_dim", 2048)):
        parser.add_argument("--" + name, type=int, default=default)
    omega.make_config(parser)
    args = parser.parse_args(argv)
    args.weight_decay = 0.05                                                          # CIFAR100.py:191-192
    args.grad_max_norm = 5.0
    model = Model(args)
    omega.InlineExecutor(folder=args.folder).submit(model)
    return model


if __name__ == "__main__":
    main()
