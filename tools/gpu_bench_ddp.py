"""BASELINE.json configs[3]: ViT-L/16 224x224 bf16 training under data parallelism, with and without overlapping the
bucketed gradient all-reduce with the fused backward.  Run under torchrun (one process per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/gpu_bench_ddp.py

Same step as bench.py (noisy-input objective, CE ls=0.1, fused AdamW), B = 128 images per GPU, CUDA-event timing,
max over ranks.  bucket_layers >= depth means ONE all-reduce after the whole backward (no overlap)."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
import vit_pytorch_robust as V  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="vit_l_16", choices=["vit_b_16", "vit_l_16"])
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
B = args.batch
g = torch.Generator().manual_seed(77 + rank)
img = torch.randn(B, 3, 224, 224, generator=g).to(torch.bfloat16).to(dev)
lab = torch.randint(0, 1000, (B,), generator=g).to(dev)
depth = {"vit_b_16": 12, "vit_l_16": 24}[args.model]
for label, buckets in (("overlapped, 3 layers per bucket", 3), ("single all-reduce after backward", depth + 1)):
    torch.manual_seed(0)
    model = getattr(V, args.model)()
    with torch.no_grad():
        model.heads.head.weight.normal_(std=0.02)
        model.class_token.normal_(std=0.02)
    model = model.to(dev)
    opt = V.FusedAdamW(model.parameters(), lr=2e-4, weight_decay=0.01)
    dp = V.DataParallel(model, optimizer=opt, bucket_layers=buckets) if world > 1 else None

    def step():
        x = img + 0.1 * torch.randn_like(img)
        opt.zero_grad()
        loss = V.softmax_cross_entropy(model(x), lab, 0.1)
        loss.backward()
        if dp is not None:
            dp.finish()
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / args.steps
    if rank == 0:
        nb = len(dp.ranges) // (args.steps + args.warmup) if dp is not None else 0
        print("%s x%d GPUs, B=%d/GPU, %-34s %7.2f ms/step  %8.0f img/s  (%d all-reduces per step, loss %.4f)" %
              (args.model, world, B, label + ":", ms, B * world / ms * 1e3, nb, loss.item()), flush=True)
    del model, opt, dp
    torch.cuda.empty_cache()
    if world == 1:
        break
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
