"""DIET / random-label training (examples/simpler_randomlabel.py:183-220 ; randomlabel.py): the backbone's classifier is
replaced by Identity (omega.utils.load_without_classifier, :127), an online probe `classifier` reads the DETACHED features,
and `extra_classifier` predicts the index of the training sample with CE(label_smoothing, executor.sh:20 uses 0.8).
Optional second parameter group for the index classifier (:255-275, --lr-scaling / --wd-scaling).

  python examples/randomlabel.py --architecture vit_b_16 --dataset imagenet --train-samples 10000 --label-smoothing 0.8
"""
import argparse

import torch

import omega_min as omega
from omega_min import V


class Model(omega.Trainer):
    def initialize_train_loader(self):
        self.num_classes, self.image_size = omega.NAME_TO_CLASS[self.args.dataset]
        if "vit" in self.args.architecture and self.args.architecture != "vit_tiny_test":
            self.image_size = 224
        self.train_samples = self.args.train_samples
        self.index_to_class = torch.arange(self.train_samples, device=self.this_device)   # simpler_randomlabel.py:186
        return omega.synthetic_loader(self.args.steps_per_epoch, self.args.batch_size // self.args.world_size,
                                      self.image_size, self.num_classes, self.this_device, seed=self.rank, with_index=True,
                                      train_samples=self.train_samples)

    def initialize_modules(self):
        model, fan_in = omega.load_without_classifier(self.args.architecture)        # :127
        self.extra_classifier = torch.nn.Linear(fan_in, self.train_samples)          # :137-141 (projector_depth 0)
        self.classifier = torch.nn.Linear(fan_in, self.num_classes)                  # :179
        self.model = model

    def initialize_optimizer(self):                                                  # :250-290
        a = self.args
        if a.lr_scaling == 1.0 and a.wd_scaling == 1.0:
            return super().initialize_optimizer()
        params = list(self.parameters())
        W = self.extra_classifier.weight
        params = [p for p in params if p is not W]
        cls = V.FusedAdamW if a.fused_optimizer else torch.optim.AdamW
        return cls([{"params": params, "lr": a.learning_rate, "weight_decay": a.weight_decay},
                    {"params": [W], "lr": a.learning_rate * a.lr_scaling, "weight_decay": a.weight_decay * a.wd_scaling}],
                   eps=1e-8, betas=(a.beta1, a.beta2))

    def compute_loss(self):                                                          # :183-220
        x = self.data[0]
        labels, indices = self.data[1].unbind(1)
        indices = self.index_to_class[indices]
        preds = self.model(x)
        preds_true = self.classifier(preds.detach())                                 # online probe
        true_loss = torch.nn.functional.cross_entropy(preds_true, labels)
        preds_false = self.extra_classifier(preds)                                   # DIET
        other_loss = V.softmax_cross_entropy(preds_false, indices, self.args.label_smoothing)
        return other_loss + true_loss

    def after_train_step(self):
        self.scheduler.step()


def main(argv=None):
    parser = argparse.ArgumentParser(description="DIET (sample-index labels) on a hot-path ViT")
    parser.add_argument("--label-smoothing", type=float, default=0.8)
    parser.add_argument("--train-samples", type=int, default=10000)                  # executor_grouped.sh:33
    parser.add_argument("--lr-scaling", type=float, default=1.0)
    parser.add_argument("--wd-scaling", type=float, default=1.0)
    omega.make_config(parser)
    args = parser.parse_args(argv)
    model = Model(args)
    omega.InlineExecutor(folder=args.folder).submit(model)
    return model


if __name__ == "__main__":
    main()
