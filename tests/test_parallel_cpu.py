"""Host-side logic of the data-parallel path, on CPU with the gloo backend and world_size 2:
bucket ranges tile the flat gradient buffer exactly once, in completion order, and the
all-reduced gradient equals the sum over ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vit_pytorch_robust as V
from vit_pytorch_robust.parallel import DataParallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeEngine:
    """Engine stand-in holding only what DataParallel touches (layout + a CPU flat gradient)."""

    def __init__(self, model):
        eng = model._nrv
        self.order, self.slots, total = eng.plan_layout()
        self.flat_grad = torch.zeros(total)
        self.flat_param = None
        self.spec = eng.spec
        self.ddp = None


def _worker(rank, world, port, bucket_layers, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = V.VisionTransformer(image_size=32, patch_size=8, num_layers=5, num_heads=2, hidden_dim=64,
                                    mlp_dim=128, num_classes=10)
        fake = _FakeEngine(model)
        model._nrv_real = model._nrv
        model._nrv = fake
        with torch.no_grad():
            for p in model.parameters():
                p.add_(rank)              # ranks start different: the constructor must broadcast rank 0's
        dp = DataParallel(model, optimizer=None, bucket_layers=bucket_layers)
        p0 = next(model.parameters()).detach().clone()
        L = fake.spec["depth"]
        g = torch.Generator().manual_seed(100 + rank)
        local = torch.randn(fake.flat_grad.numel(), generator=g)
        # emulate the backward: a stage's gradients become final only when that stage has run
        fake.flat_grad.zero_()
        chunks = dp.stage_chunks(L)
        written = 0
        for hi, lo in chunks:
            end = dp._range_end_for_stage(fake, lo)
            fake.flat_grad[written:end] = local[written:end]
            written = end
            dp.stages_done(fake, hi, lo)
        dp.finish()
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        want = sum(gathered) / world
        q.put((rank, chunks, dp.ranges, bool(torch.allclose(fake.flat_grad, want, atol=1e-6)),
               float(p0.flatten()[0]), fake.flat_grad.numel()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bucket_layers", [1, 2, 7])
def test_bucketed_allreduce_gloo_world2(bucket_layers):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, bucket_layers, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    res.sort()
    for rank, chunks, ranges, ok, p0, total in res:
        assert ok, "all-reduced gradient != mean over ranks"
        # stages: contiguous, descending, from depth down to -1
        assert chunks[0][0] == 5 and chunks[-1][1] == -1
        for (h1, l1), (h2, l2) in zip(chunks, chunks[1:]):
            assert h2 == l1 - 1 and l1 <= h1 and l2 <= h2
        # buckets tile [0, total) exactly once, front to back
        assert ranges[0][0] == 0 and ranges[-1][1] == total
        for (s1, e1), (s2, e2) in zip(ranges, ranges[1:]):
            assert e1 == s2 and s1 < e1
    assert res[0][4] == res[1][4]  # parameters were broadcast from rank 0


def test_stage_chunks_cover_all_stages():
    class E:  # minimal engine
        ddp = None
        flat_param = None
    m = type("M", (), {"_nrv": E(), "parameters": lambda self: iter(())})()
    dp = DataParallel.__new__(DataParallel)
    for L in (1, 2, 6, 12, 24):
        for bl in (1, 2, 3, 5, 100):
            dp.bucket_layers = bl
            chunks = dp.stage_chunks(L)
            stages = [s for hi, lo in chunks for s in range(hi, lo - 1, -1)]
            assert stages == list(range(L, -2, -1)), (L, bl, chunks)
