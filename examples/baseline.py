"""Supervised baseline and the noisy-input objective on a hot-path VisionTransformer.

baseline.py:63-71 of the reference: preds = model(x); CE(label_smoothing=0.1); AdamW; grad_max_norm 5.0 (:121).
--noise-std > 0 selects the objective of examples/nowak.py:148-158: the batch [x + eps, x] goes through the network and
the loss is taken on the noisy half (or on the clean half with --improved).  For a ViT (no batch statistics) only the half
the loss reads matters, so one half is computed; eps comes from nrv_add_gaussian_noise.

  python examples/baseline.py --architecture vit_b_16 --dataset imagenet --batch-size 256 --noise-std 0.1
"""
import argparse

import torch

import omega_min as omega
from omega_min import V


class Model(omega.Trainer):
    def initialize_train_loader(self):
        self.num_classes, self.image_size = omega.NAME_TO_CLASS[self.args.dataset]
        if "vit" in self.args.architecture and self.args.architecture != "vit_tiny_test":
            self.image_size = 224                                                     # sup_ssl.py:26-27
        return omega.synthetic_loader(self.args.steps_per_epoch, self.args.batch_size // self.args.world_size,
                                      self.image_size, self.num_classes, self.this_device, seed=self.rank)

    def initialize_val_loader(self):
        return omega.synthetic_loader(2, self.args.batch_size // self.args.world_size, self.image_size, self.num_classes,
                                      self.this_device, seed=1000 + self.rank)

    def initialize_modules(self):
        model, fan_in = omega.load_without_classifier(self.args.architecture)
        model.heads.head = torch.nn.Linear(fan_in, self.num_classes)                 # baseline.py:61 (model.fc = Linear(...))
        with torch.no_grad():
            model.heads.head.weight.normal_(std=0.02)
        self.model = model

    def compute_loss(self):
        x, y = self.data
        if self.args.noise_std > 0 and not self.args.improved:                       # nowak.py:153-158
            x = V.add_gaussian_noise(x, self.args.noise_std)
        preds = self.model(x)
        return V.softmax_cross_entropy(preds, y, 0.1)                                # baseline.py:70

    def before_eval_epoch(self):
        super().before_eval_epoch()
        self.accu, self.counter = 0.0, 0

    def eval_step(self):                                                             # baseline.py:78-85
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
    parser = argparse.ArgumentParser(description="supervised / noisy-input training of a hot-path ViT")
    parser.add_argument("--noise-std", type=float, default=0.0)
    parser.add_argument("--improved", action="store_true")
    omega.make_config(parser)
    args = parser.parse_args(argv)
    args.grad_max_norm = 5.0
    model = Model(args)
    omega.InlineExecutor(folder=args.folder).submit(model)
    return model


if __name__ == "__main__":
    main()
