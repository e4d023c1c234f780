"""bench.py on a box without a GPU: the reference arm prints the contract's JSON line (the reference's CPU path timed on
the host cores), and the product arm refuses to run instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="" if not torch.cuda.is_available() else os.environ.get("CUDA_VISIBLE_DEVICES", ""))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT, env=env)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "train_images_per_sec" and line["unit"] == "images/s"
    assert line["higher_is_better"] is True and line["steps"] == 1 and line["n_gpus"] == 1
    assert line["value"] > 0 and abs(line["value"] - 8 * 1e3 / line["ms_per_step"]) < 1e-6 * line["value"] + 1e-9
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU")
def test_product_arm_refuses_to_run_without_a_gpu():
    r = _run("--steps", "1", "--warmup", "1", timeout=300)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
