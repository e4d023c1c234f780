#!/usr/bin/env python
"""Headline benchmark: ViT-B/16 224x224 BF16 training throughput (images/s) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path over one batch of synthetic images (BASELINE.json configs[2]):
noisy-input objective (x + 0.1*randn, examples/nowak.py:153) -> VisionTransformer forward ->
softmax cross-entropy (label smoothing 0.1) -> backward -> gradient all-reduce (N > 1) -> AdamW
(lr 2e-4, wd 0.01; examples/executor.sh:16-21).  Prints ONE JSON line (see the contract in the
task statement); `value` is device-resident throughput, `e2e` includes pinned-host -> device
copies of every batch and a device -> host read of every step's loss.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "noise-robust-vit_b200"))

IMG, PATCH, DIM, DEPTH, HEADS, MLP, CLASSES = 224, 16, 768, 12, 12, 3072, 1000
# --model: the headline is ViT-B/16 (BASELINE.json configs[2]); l16 = configs[3] (ViT-L/16, B=128 per GPU)
MODELS = {"b16": dict(patch=16, D=768, L=12, H=12, M=3072, batch=256, name="ViT-B/16", ctor="vit_b_16", cfg="configs[2]"),
          "l16": dict(patch=16, D=1024, L=24, H=16, M=4096, batch=128, name="ViT-L/16", ctor="vit_l_16", cfg="configs[3]")}


def flops_per_image_fwd(img=IMG, patch=PATCH, D=DIM, L=DEPTH, H=HEADS, M=MLP, C=CLASSES, cls=1, ch=3):
    """F = 2 n (ch P^2) D + L [2 N D 3I + 4 N^2 I + 2 N I D + 4 N D M] + 2 D C."""
    n = (img // patch) ** 2
    N = n + cls
    inner = D
    return 2 * n * (ch * patch * patch) * D + L * (2 * N * D * 3 * inner + 4 * N * N * inner + 2 * N * inner * D + 4 * N * D * M) + 2 * D * C


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return p, "measured"
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            # nvidia-smi initialises NVML for a few hundred ms and holds driver locks while it does: with a short warm-up
            # that landed INSIDE the timed region and starved the launch thread (one run in six lost 25 %, SM clocks at
            # idle values).  Wait for its first line before any step is issued.
            t_end = time.time() + 5.0
            while not self.lines and time.time() < t_end and self.proc.poll() is None:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Clocks over the samples that arrived inside [t0, t1] (the timed region).  nvidia-smi takes a few hundred
        ms to deliver its first line, so the sampler is started before the warm-up; when the timed region is
        too short to hold a sample, the warm-up samples (same load) are used and `window` says so."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        lines = [ln for (t, ln) in self.lines if t0 is None or (t0 <= t <= t1)]
        window = "timed region"
        if not lines:
            lines = [ln for (t, ln) in self.lines if t <= (t1 or t)]
            window = "warm-up + timed region (timed region shorter than the sampling latency)"
        sm, mx, reasons = [], None, set()
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median over the samples taken under load (upper half of the distribution when idle samples sneak in)
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "window": window}


def cpu_reference_step_fn(batch, threads):
    """The reference's own CPU path for this workload, as BASELINE.md section 4 defines it: for vit.py models the
    state-dict-identical torchvision VisionTransformer (the class vit.py:178-351 was copied from), because the reference's
    VisionTransformer.forward raises as shipped (utils.py:877, utils.py:210) and /root/reference does not travel to the GPU
    box.  fp32, all host threads, noisy-input objective + CE(ls 0.1) + backward.  Returns (step, kind, what); falls back to
    the oracle restatement (kind "port") when torchvision is not importable."""
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    img = torch.randn(batch, 3, IMG, IMG, generator=g)
    labels = torch.randint(0, CLASSES, (batch,), generator=g)
    try:
        from torchvision.models.vision_transformer import vit_b_16 as tv_vit_b_16
        torch.manual_seed(0)
        twin = tv_vit_b_16()
        with torch.no_grad():
            twin.heads.head.weight.normal_(std=0.02)
        twin.train()

        def step():
            x = img + 0.1 * torch.randn_like(img)
            for p in twin.parameters():
                p.grad = None
            loss = torch.nn.functional.cross_entropy(twin(x), labels, label_smoothing=0.1)
            loss.backward()
            return loss.item()

        return step, "reference", "torchvision VisionTransformer vit_b_16 (state-dict-identical twin of vit.py:377-403, BASELINE.md 4)"
    except Exception:   # noqa: BLE001 - no torchvision on this host
        pass
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vit_oracle as O
    import vit_pytorch_robust as V
    torch.manual_seed(0)
    shell = V.vit_b_16()  # parameter container only (CPU tensors, never executed)
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in shell.state_dict().items()}
    with torch.no_grad():
        sd["heads.head.weight"].normal_(std=0.02)
    params = [v for v in sd.values() if v.requires_grad]

    def step():
        x = img + 0.1 * torch.randn_like(img)
        logits = O.vision_transformer_forward(sd, x, patch_size=PATCH, num_heads=HEADS)
        loss = O.cross_entropy(logits, labels, 0.1)
        torch.autograd.grad(loss, params)
        return loss.item()

    return step, "port", "oracle restatement (oracle/vit_oracle.py) of the reference module"


def time_cpu_baseline(batch=8, warmup=1, iters=3):
    threads = os.cpu_count() or 1
    step, kind, what = cpu_reference_step_fn(batch, threads)
    for _ in range(warmup):
        step()
    best = None
    for _ in range(iters):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": batch / best, "unit": "images/s", "cores": threads, "kind": kind,
            "sample": "%s: ViT-B/16 224^2 fp32 noisy-input fwd+CE+bwd, batch %d, best of %d steps after %d warm-up" %
                      (what, batch, iters, warmup)}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    batch = 8
    threads = os.cpu_count() or 1
    step, kind, what = cpu_reference_step_fn(batch, threads)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = batch * args.steps / dt
    line = {
        "impl": "reference", "metric": "train_images_per_sec", "value": val, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ViT-B/16 224x224 training step (noisy-input objective, CE ls=0.1), CPU, batch %d per step" % batch},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": kind,
                         "sample": "%s on host cores, batch %d x %d steps" % (what, batch, args.steps)},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    # defaults: ~0.7 s of warm-up so the power-cap controller has settled before the timed region (with 5 warm-up
    # steps the first timed steps ran 2-3 % slower than the end-to-end region measured later in the same process)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step (default: 256 for b16, 128 for l16)")
    ap.add_argument("--model", default="b16", choices=sorted(MODELS), help="b16: the headline config; l16: BASELINE.json configs[3]")
    ap.add_argument("--robust", action="store_true", help="robust=True: Sinkhorn attention (utils.py:1025-1037) in every layer")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--attn-impl", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--ln-mode", default="default", choices=["default", "folded", "separate"],
                    help="LayerNorm in front of the QKV / FC1 GEMMs: folded into them, or stand-alone kernels (training default)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    spec = MODELS[args.model]
    if args.batch is None:
        args.batch = spec["batch"]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return 0

    import torch.distributed as dist
    import vit_pytorch_robust as V
    from vit_pytorch_robust import _abi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    torch.manual_seed(0)
    model = getattr(V, spec["ctor"])(robust=args.robust)
    with torch.no_grad():  # the reference zero-initialises the head (vit.py:304-306): give the loss a gradient
        model.heads.head.weight.normal_(std=0.02)
        model.class_token.normal_(std=0.02)
    model = model.to(dev)
    model._nrv.attn_impl = {"auto": _abi.ATTN_IMPL_AUTO, "simt": _abi.ATTN_IMPL_SIMT, "tc": _abi.ATTN_IMPL_TC}[args.attn_impl]
    if args.ln_mode != "default":
        model._nrv.ln_mode_train = _abi.LN_FOLDED if args.ln_mode == "folded" else _abi.LN_SEPARATE
    opt = V.FusedAdamW(model.parameters(), lr=2e-4, weight_decay=0.01)
    dp = V.DataParallel(model, optimizer=opt, bucket_layers=3) if world > 1 else None

    g = torch.Generator().manual_seed(1234 + rank)
    host_img = [torch.randn(B, 3, IMG, IMG, generator=g).to(torch.bfloat16).pin_memory() for _ in range(2)]
    host_lab = [torch.randint(0, CLASSES, (B,), generator=g).pin_memory() for _ in range(2)]
    dev_img = host_img[0].to(dev)
    dev_lab = host_lab[0].to(dev)
    lib = _abi.load()

    def train_step(img, labels):
        x = V.add_gaussian_noise(img, 0.1)               # noisy-input objective (nowak.py:153), one libnrvit kernel
        opt.zero_grad()
        logits = model(x)
        loss = V.softmax_cross_entropy(logits, labels, 0.1)
        loss.backward()
        if dp is not None:
            dp.finish()
        opt.step()
        return loss

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- device-resident timing (`value`)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # Settle phase (disclosed in config.settle_steps): with a short warm-up (the driver's --warmup 5 = 0.16 s) the power-cap
    # controller is still hunting when the timed region starts -- the boxes of this pool start a fresh process at 1.9 GHz,
    # overshoot the 1 kW cap and undershoot for a few hundred ms; 3 of 25 such runs measured 7-10 % low while the later
    # regions of the same process were normal.  Untimed steps are topped up to 20 (~0.65 s) before the W warm-up steps.
    settle_steps = max(0, 20 - args.warmup)
    for _ in range(settle_steps):
        train_step(dev_img, dev_lab)
    for _ in range(args.warmup):
        train_step(dev_img, dev_lab)
    sync_all()
    t_wall0 = time.time()
    l0 = lib.nrv_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = train_step(dev_img, dev_lab)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    t_wall1 = time.time()
    launches = lib.nrv_launch_count() - l0
    # Roofline pass: the SAME K steps again, now with a CUDA-event pair around every GEMM launch (on the stream
    # the kernels run on).  Kept out of the headline region because ~300 event records per step cost it 1-3 %.
    import ctypes as C
    lib.nrv_gemm_timing(1)
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(args.steps):
        loss = train_step(dev_img, dev_lab)
    r1.record()
    sync_all()
    ms_roof = r0.elapsed_time(r1)
    g_ms, g_fl, g_n = C.c_double(), C.c_double(), C.c_longlong()
    lib.nrv_gemm_timing_read(C.byref(g_ms), C.byref(g_fl), C.byref(g_n))
    lib.nrv_gemm_timing(0)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    final_loss = loss.item()

    # ---------------- end-to-end timing (`e2e`): pinned host -> device every step, loss read back
    copy_stream = torch.cuda.Stream(device=dev)
    dbuf = [torch.empty_like(dev_img) for _ in range(2)]
    lbuf = [torch.empty_like(dev_lab) for _ in range(2)]
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def issue_copy(i):
        s = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])
            dbuf[s].copy_(host_img[s], non_blocking=True)
            lbuf[s].copy_(host_lab[s], non_blocking=True)
            ready[s].record(copy_stream)

    def e2e_loop(n):
        cur = torch.cuda.current_stream()
        for s in range(2):
            consumed[s].record(cur)
        issue_copy(0)
        seen = 0.0
        for i in range(n):
            s = i & 1
            if i + 1 < n:
                issue_copy(i + 1)                         # prefetch the next batch under this step's compute
            cur.wait_event(ready[s])
            loss = train_step(dbuf[s], lbuf[s])
            consumed[s].record(cur)
            loss_host[s].copy_(loss.detach(), non_blocking=True)
            if i > 0:
                seen += float(loss_host[(i - 1) & 1])      # previous step's loss (already landed or nearly so)
        torch.cuda.synchronize()
        seen += float(loss_host[(n - 1) & 1])
        return seen

    e2e_loop(3)
    sync_all()
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = t.item()

    if rank == 0:
        peaks, peak_src = measured_peaks()
        imgs = B * world * args.steps
        value = imgs / (ms / 1e3)
        train_flops = 3 * flops_per_image_fwd(patch=spec["patch"], D=spec["D"], L=spec["L"], H=spec["H"], M=spec["M"])
        peak_tf = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        ach = (g_fl.value / 1e12) / (g_ms.value / 1e3) if g_ms.value > 0 else 0.0
        prof = {}
        pj = os.path.join(ROOT, "profiles", "ncu_top_kernel.json")
        if os.path.exists(pj):
            with open(pj) as fh:
                prof = json.load(fh)
        line = {
            "metric": "train_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "%s 224x224 training step%s: noisy-input objective + CE(ls=0.1) + AdamW (BASELINE.json %s)" %
                                   (spec["name"], " (robust=True: Sinkhorn attention)" if args.robust else "", spec["cfg"]),
                       "per_gpu_batch": B, "global_batch": B * world, "tokens": 197, "parallelism": "dp%d" % world,
                       "settle_steps": settle_steps,   # untimed steps in front of the W warm-up steps (power-cap settling)
                       "l2_policy": "inputs+activations per step (>10 GB) exceed the 126 MB L2; no flush needed",
                       "attention": args.attn_impl,
                       "layernorm": "folded into the QKV / FC1 GEMMs" if model._nrv.ln_mode_train == _abi.LN_FOLDED else "stand-alone kernels",
                       "model_tflops": value * train_flops / 1e12,
                       "frac_of_sustained_bf16_peak": value * train_flops / 1e12 / peaks.get("bf16_tflops_sustained", 1393.0),
                       "frac_of_burst_bf16_peak": value * train_flops / 1e12 / peaks["bf16_tflops"],
                       "final_loss": final_loss},
            "roofline": {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": ach / peak_tf if peak_tf else None, "traffic": prof.get("dram_bytes_per_launch"),
                         "kernel": "nrv::gemm_kernel (tcgen05 GEMM; CUDA-event pairs around all %d launches of a second "
                                   "pass over the same K steps)" % g_n.value,
                         "peak_source": "%s (bf16_tflops_sustained: kernel timed inside a long step)" % peak_src,
                         "gemm_share_of_step": (g_ms.value / ms_roof) if ms_roof > 0 else None,
                         "ms_per_step_with_events": ms_roof / args.steps},
            "clocks": clocks,
            "e2e": {"value": imgs / (ms_e2e / 1e3), "unit": "images/s",
                    "h2d_bytes_per_step": host_img[0].numel() * 2 + host_lab[0].numel() * 8, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches),
        }
        if world == 1 and not args.no_cpu_baseline and args.model == "b16" and not args.robust:
            line["cpu_baseline"] = time_cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        if dp is not None:
            dp.close()            # the library's NCCL communicator goes before the process group and the CUDA context
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
