"""The example training scripts (examples/*.py: the training loops of the reference's examples/CIFAR100.py, baseline.py /
nowak.py, simpler_randomlabel.py and sup_ssl.py restated on the hot-path models over a minimal omega-style Trainer) run a few
steps on the B200: losses finite, parameters move, the loss of a repeated batch goes down."""
import importlib
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

EX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "examples")
if EX not in sys.path:
    sys.path.insert(0, EX)

TINY = ["--architecture", "vit_tiny_test", "--dataset", "cifar100", "--batch-size", "32", "--steps-per-epoch", "1",
        "--epochs", "6", "--learning-rate", "1e-3"]


def _run(name, argv):
    torch.manual_seed(0)
    mod = importlib.import_module(name)
    trainer = mod.main(argv)
    losses = trainer.losses
    assert all(l == l and abs(l) < 1e6 for l in losses), losses
    return trainer, losses


def test_cifar100_simplevit_script_trains():
    t, losses = _run("cifar100_simplevit", ["--dataset", "cifar100", "--batch-size", "64", "--steps-per-epoch", "1",
                                            "--epochs", "12", "--dim", "128", "--depth", "2", "--heads", "2", "--mlp_dim", "256",
                                            "--learning-rate", "2e-3", "--eval-each-epoch"])
    assert losses[-1] < losses[0] - 0.05                      # one batch repeated 12 times: it is being fitted
    assert any(name == "eval_accuracies" for name, _ in t.logs)
    assert t.args.grad_max_norm == 5.0 and t.args.weight_decay == 0.05


def test_cifar100_simplevit_script_with_cutmix_and_sinkhorn_attention():
    _run("cifar100_simplevit", ["--dataset", "cifar100", "--batch-size", "32", "--steps-per-epoch", "2", "--epochs", "2",
                                "--dim", "128", "--depth", "2", "--heads", "2", "--mlp_dim", "256", "--cutmix_prob", "1.0", "--robust",
                                "--fused-optimizer"])


@pytest.mark.parametrize("extra", [[], ["--noise-std", "0.1"], ["--noise-std", "0.1", "--improved"]])
def test_baseline_and_noisy_input_script_trains(extra):
    t, losses = _run("baseline", TINY + extra)
    assert losses[-1] < losses[0]


def test_randomlabel_diet_script_trains_with_two_parameter_groups():
    t, losses = _run("randomlabel", TINY + ["--train-samples", "512", "--label-smoothing", "0.8", "--lr-scaling", "2.0",
                                            "--fused-optimizer"])
    assert losses[-1] < losses[0]
    assert len(t.optimizer.param_groups) == 2
    assert isinstance(t.model.heads.head, torch.nn.Identity)  # features feed classifier (detached) and extra_classifier


def test_sup_ssl_script_trains():
    t, losses = _run("sup_ssl", TINY + ["--temperature", "0.5"])
    assert losses[-1] < losses[0]
