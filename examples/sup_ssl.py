"""Supervised-SSL objective of examples/sup_ssl.py:100-124 on a hot-path ViT: a VICReg-style loss on a linear projector of the
features -- covariance towards the identity plus invariance between samples of the same class -- next to an online probe on
the detached features.  MultiStepLR at 50 % / 75 % of training (:148-159).

  python examples/sup_ssl.py --architecture vit_b_16 --dataset imagenet --temperature 1.0
"""
import argparse

import torch

import omega_min as omega


class Model(omega.Trainer):
    def initialize_train_loader(self):
        self.num_classes, self.image_size = omega.NAME_TO_CLASS[self.args.dataset]   # sup_ssl.py:24-27
        if "vit" in self.args.architecture and self.args.architecture != "vit_tiny_test":
            self.image_size = 224
        return omega.synthetic_loader(self.args.steps_per_epoch, self.args.batch_size // self.args.world_size,
                                      self.image_size, self.num_classes, self.this_device, seed=self.rank, with_index=True,
                                      train_samples=1 << 20)

    def initialize_modules(self):                                                    # sup_ssl.py:76-104
        model, fan_in = omega.load_without_classifier(self.args.architecture)
        self.projector = torch.nn.Linear(fan_in, self.num_classes)
        self.classifier = torch.nn.Linear(fan_in, self.num_classes)
        self.model = model

    def compute_loss(self):                                                          # sup_ssl.py:106-124
        x = self.data[0]
        labels = self.data[1][:, 0]
        preds = self.model(x)
        preds_true = self.classifier(preds.detach())
        true_loss = torch.nn.functional.cross_entropy(preds_true, labels)
        G = labels[:, None].eq(labels)
        Z = self.projector(preds)
        Cm = torch.cov(Z.t())
        eye = torch.eye(Cm.size(0), dtype=Cm.dtype, device=Cm.device)
        VC_loss = (Cm - eye).square().mean()
        i, j = G.nonzero(as_tuple=True)
        inv_loss = (Z[i] - Z[j]).square().mean()
        return VC_loss + self.args.temperature * inv_loss + true_loss

    def initialize_scheduler(self):                                                  # sup_ssl.py:148-159
        N = len(self.train_loader)
        return torch.optim.lr_scheduler.MultiStepLR(
            self.optimizer, milestones=[int(self.args.epochs * 0.5) * N, int(self.args.epochs * 0.75) * N], gamma=0.1)

    def after_train_step(self):
        self.scheduler.step()


def main(argv=None):
    parser = argparse.ArgumentParser(description="supervised-SSL objective on a hot-path ViT")
    parser.add_argument("--temperature", type=float, default=1.0)
    omega.make_config(parser)
    args = parser.parse_args(argv)
    model = Model(args)
    omega.InlineExecutor(folder=args.folder).submit(model)
    return model


if __name__ == "__main__":
    main()
