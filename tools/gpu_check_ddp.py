"""Multi-GPU check (run under torchrun, 1 process per GPU): bucketed, overlapped all-reduce inside the
fused backward gives the same gradients / parameters as one process on the concatenated batch."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "noise-robust-vit_b200"))
import vit_pytorch_robust as V  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = dict(image_size=64, patch_size=16, num_layers=4, num_heads=4, hidden_dim=256, mlp_dim=512, num_classes=24)
    B = 8
    g = torch.Generator().manual_seed(7)
    img_all = torch.randn(B * world, 3, 64, 64, generator=g)
    lab_all = torch.randint(0, 24, (B * world,), generator=g)
    ok = True
    for mode in (torch.float32, torch.bfloat16):
        torch.manual_seed(1)
        model = V.VisionTransformer(**cfg)
        with torch.no_grad():
            model.heads.head.weight.normal_(std=0.05)
            model.class_token.normal_(std=0.05)
        ref = V.VisionTransformer(**cfg)
        ref.load_state_dict(model.state_dict())
        model, ref = model.to(dev), ref.to(dev)
        model._nrv.compute_dtype = mode
        ref._nrv.compute_dtype = mode
        opt = V.FusedAdamW(model.parameters(), lr=1e-3, weight_decay=0.01)
        ropt = V.FusedAdamW(ref.parameters(), lr=1e-3, weight_decay=0.01)
        dp = V.DataParallel(model, optimizer=opt, bucket_layers=1)
        for step in range(2):
            opt.zero_grad()
            x = img_all[rank * B:(rank + 1) * B].to(dev)
            y = lab_all[rank * B:(rank + 1) * B].to(dev)
            V.softmax_cross_entropy(model(x), y, 0.1).backward()
            dp.finish()
            ropt.zero_grad()
            V.softmax_cross_entropy(ref(img_all.to(dev)), lab_all.to(dev), 0.1).backward()
            torch.cuda.synchronize()
            # summed per-rank mean-loss gradients / world == full-batch mean-loss gradient
            worst = 0.0
            for (k, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
                worst = max(worst, rel(p.grad / world, q.grad))
            tol = 2e-4 if mode == torch.float32 else 3e-2
            good = worst < tol
            ok &= good
            if rank == 0:
                print("mode %s step %d: worst grad rel %.3e (tol %.0e) buckets %d %s" %
                      (mode, step, worst, tol, len(dp.ranges), "ok" if good else "FAIL"), flush=True)
            dp.ranges.clear()
            opt.step()
            ropt.step()
        torch.cuda.synchronize()
        # Adam turns noise-level gradient differences of near-zero-gradient elements into +-lr steps, so
        # compare the whole parameter vector (tiny bias tensors would dominate a per-tensor maximum)
        worst = rel(model._nrv.flat_param, ref._nrv.flat_param)
        good = worst < (1e-4 if mode == torch.float32 else 2e-3)
        ok &= good
        if rank == 0:
            print("mode %s: parameters after 2 steps rel %.3e %s" % (mode, worst, "ok" if good else "FAIL"), flush=True)
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
