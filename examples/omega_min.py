"""Minimal stand-in for the author's private `omega` trainer, so that the training loops of the reference's example scripts
can run on the B200-native models (test / bench harness, not product).

The reference scripts subclass `omega.Trainer` and fill in hooks (examples/CIFAR100.py:16-160, baseline.py:8-97,
simpler_randomlabel.py:20-290, sup_ssl.py:20-160, nowak.py:48-180).  `omega`, `submitit`, `ffcv` and `torchmetrics` are not
installable here and the datasets live on the author's cluster, so this module restates the part of the contract those
scripts rely on, reconstructed from their call sites:

  Trainer(args)                      .args, .this_device, .model, .optimizer, .scheduler, .train_loader, .val_loader, .data
  hooks                              initialize_train_loader / initialize_val_loader / initialize_modules /
                                     initialize_optimizer / initialize_scheduler / compute_loss / before_train_step /
                                     after_train_step / before_eval_epoch / eval_step / after_eval_epoch / log_txt
  trainer()                          epochs x (train steps with backward, grad_max_norm clipping, optimizer step;
                                     optional evaluation epoch)
  make_config(parser)                the flags `omega.argparse.make_config` adds (executor.sh:16-21 values as defaults)
  load_without_classifier(arch)      (model with its classifier replaced by nn.Identity, fan_in); vit_* route to
                                     vit_pytorch_robust.vit (evaluation.py:129-131 does the same by hand)
  synthetic_loader(...)              random images / labels of the dataset's shape (there are no datasets on the box)

Multi-GPU: one process per GPU under torchrun; gradients of the fused encoder go through vit_pytorch_robust.DataParallel.
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "noise-robust-vit_b200"))

import vit_pytorch_robust as V  # noqa: E402

# name -> (num_classes, image_size), as omega.dataset.NAME_TO_CLASS is used in sup_ssl.py:24-25
NAME_TO_CLASS = {"cifar10": (10, 32), "cifar100": (100, 32), "tinyimagenet": (200, 64), "imagenet": (1000, 224)}


def make_config(parser):
    """Flags the reference scripts read from `self.args` after `omega.argparse.make_config(parser)`."""
    existing = {a.dest for a in parser._actions}

    def add(flag, **kw):
        if flag.lstrip("-").replace("-", "_") not in existing:
            parser.add_argument(flag, **kw)
    add("--architecture", type=str, default="vit_b_16")
    add("--batch-size", type=int, default=256)
    add("--epochs", type=int, default=1)
    add("--learning-rate", type=float, default=2e-4)       # executor.sh:16
    add("--weight-decay", type=float, default=0.01)        # executor.sh:17
    add("--beta1", type=float, default=0.9)
    add("--beta2", type=float, default=0.999)
    add("--grad-max-norm", type=float, default=None)
    add("--eval-each-epoch", action="store_true")
    add("--float16", action="store_true")
    add("--folder", type=str, default="./omega_min_out")
    add("--steps-per-epoch", type=int, default=4, help="length of the synthetic loader")
    add("--dataset", type=str, default="cifar100", choices=sorted(NAME_TO_CLASS))
    add("--fused-optimizer", action="store_true", help="FusedAdamW (nrv_adamw) instead of torch.optim.AdamW")
    return parser


def synthetic_loader(steps, batch, image_size, num_classes, device, seed=0, with_index=False, train_samples=None):
    """`steps` batches of (images [B,3,S,S] fp32, labels int64 [B]) resident on `device`; with_index: labels [B, 2] =
    (class, sample index) as the ffcv pipelines of simpler_randomlabel.py deliver them (compute_loss :185)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(steps):
        x = torch.randn(batch, 3, image_size, image_size, generator=g)
        y = torch.randint(0, num_classes, (batch,), generator=g)
        if with_index:
            idx = torch.randint(0, train_samples, (batch,), generator=g)
            y = torch.stack([y, idx], 1)
        out.append((x.to(device), y.to(device)))
    return out


def load_without_classifier(arch, **kw):
    """omega.utils.load_without_classifier (sup_ssl.py:90, simpler_randomlabel.py:127): backbone + feature width."""
    if arch in ("vit_b_16", "vit_b_32", "vit_l_16", "vit_l_32", "vit_h_14"):
        model = getattr(V, arch)(**kw)
        fan_in = model.heads.head.in_features
        model.heads.head = torch.nn.Identity()             # evaluation.py:129-131
        return model, fan_in
    if arch == "vit_tiny_test":                             # a few-layer VisionTransformer for the harness tests
        cfg = dict(image_size=32, patch_size=8, num_layers=2, num_heads=2, hidden_dim=128, mlp_dim=256)
        cfg.update(kw)
        model = V.VisionTransformer(**cfg)
        model.heads.head = torch.nn.Identity()
        return model, cfg["hidden_dim"]
    raise ValueError("architecture %r is outside the hot path (vit_b_16 ... vit_h_14)" % arch)


class Trainer(torch.nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.args.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("the examples run the B200-native encoder: no CUDA device (there is no CPU fallback)")
        torch.cuda.set_device(local)
        self.this_device = torch.device("cuda", local)
        self.logs = []

    # ---- hooks every script fills in
    def initialize_train_loader(self):
        raise NotImplementedError

    def initialize_modules(self):
        raise NotImplementedError

    def compute_loss(self):
        raise NotImplementedError

    # ---- defaults of the hooks the scripts may leave alone (CIFAR100.py:90-97 optimiser; omega's cosine schedule)
    def initialize_optimizer(self):
        cls = V.FusedAdamW if getattr(self.args, "fused_optimizer", False) else torch.optim.AdamW
        return cls(self.parameters(), lr=self.args.learning_rate, weight_decay=self.args.weight_decay, eps=1e-8,
                   betas=(self.args.beta1, self.args.beta2))

    def initialize_scheduler(self):
        return torch.optim.lr_scheduler.CosineAnnealingLR(self.optimizer, T_max=max(1, self.args.epochs * len(self.train_loader)))

    def initialize_val_loader(self):
        return None

    def before_train_step(self):
        pass

    def after_train_step(self):
        pass

    def before_eval_epoch(self):
        self.eval()

    def after_eval_epoch(self):
        self.train()

    def eval_step(self):
        pass

    def log_txt(self, name, **kw):
        self.logs.append((name, kw))
        if self.rank == 0:
            print("[%s] %s" % (name, " ".join("%s=%.5g" % (k, v) for k, v in kw.items())), flush=True)

    # ---- the loop
    def __call__(self):
        a = self.args
        if a.world_size > 1 and not torch.distributed.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            torch.distributed.init_process_group("nccl", device_id=self.this_device)
        self.train_loader = self.initialize_train_loader()
        self.val_loader = self.initialize_val_loader()
        self.initialize_modules()
        self.to(self.this_device)
        self.optimizer = self.initialize_optimizer()
        self.scheduler = self.initialize_scheduler()
        # gradient all-reduce: the fused encoder's flat buffer in buckets (overlapped with its backward), the heads trained
        # next to it (classifier / extra_classifier / projector) in finish(); averaged over the ranks as torch DDP does
        extra = [m for m in self.children() if m is not self.model]
        self.dp = V.DataParallel(self.model, optimizer=None, extra_modules=extra) if a.world_size > 1 else None
        losses = []
        t0 = time.time()
        for epoch in range(a.epochs):
            self.train()
            for self.data in self.train_loader:
                self.before_train_step()
                self.optimizer.zero_grad(set_to_none=False)
                loss = self.compute_loss()
                loss.backward()
                if self.dp is not None:
                    self.dp.finish()
                if getattr(a, "grad_max_norm", None):
                    V.clip_grad_norm_(self.parameters(), a.grad_max_norm)       # CIFAR100.py:192
                self.optimizer.step()
                self.after_train_step()
                losses.append(loss.detach())
            if a.eval_each_epoch and self.val_loader is not None:
                self.before_eval_epoch()
                with torch.no_grad():
                    for self.data in self.val_loader:
                        self.eval_step()
                self.after_eval_epoch()
        torch.cuda.synchronize()
        self.losses = [float(x) for x in losses]
        self.log_txt("train", first_loss=self.losses[0], last_loss=self.losses[-1], steps=len(self.losses),
                     seconds=time.time() - t0)
        return self.losses

class InlineExecutor:
    """submitit.AutoExecutor as the scripts use it (CIFAR100.py:196-218): here `submit` runs the trainer in-process."""

    def __init__(self, folder=None):
        self.folder = folder

    def update_parameters(self, **kw):
        self.parameters = kw

    def submit(self, fn):
        class Job:
            job_id = "inline"
            result = fn()
        return Job()
