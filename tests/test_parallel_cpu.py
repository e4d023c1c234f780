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

    def _view(self, buf, s):
        return buf[s.offset:s.offset + s.numel].view(s.param.shape)


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
    dp.sync, dp._reduced_since_zero = True, False
    for L in (1, 2, 6, 12, 24):
        for bl in (1, 2, 3, 5, 100):
            dp.bucket_layers = bl
            chunks = dp.stage_chunks(L)
            stages = [s for hi, lo in chunks for s in range(hi, lo - 1, -1)]
            assert stages == list(range(L, -2, -1)), (L, bl, chunks)


def _worker_nosync_foreign(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = V.VisionTransformer(image_size=32, patch_size=8, num_layers=3, num_heads=2, hidden_dim=64,
                                    mlp_dim=128, num_classes=10)
        fake = _FakeEngine(model)
        model._nrv = fake
        probe = torch.nn.Linear(10, 3)
        dp = DataParallel(model, optimizer=None, bucket_layers=1, extra_modules=[probe])
        L = fake.spec["depth"]
        g = torch.Generator().manual_seed(100 + rank)
        mb1 = torch.randn(fake.flat_grad.numel(), generator=g)
        mb2 = torch.randn(fake.flat_grad.numel(), generator=g)
        fg = [torch.randn(p.shape, generator=g) for p in probe.parameters()]
        fake.flat_grad.zero_()
        dp.on_zero_grad()
        # micro-batch 1 under no_sync(): one chunk, no collective, nothing marked as reduced
        with dp.no_sync():
            chunks = dp.stage_chunks(L)
            assert chunks == [(L, -1)]
            fake.flat_grad += mb1
            dp.stages_done(fake, L, -1)
            dp.finish()
        assert not dp.ranges and torch.equal(fake.flat_grad, mb1)
        # micro-batch 2 synchronised: the ACCUMULATED gradient is reduced exactly once
        for p, gr in zip(probe.parameters(), fg):
            p.grad = gr.clone()
        written = 0
        for hi, lo in dp.stage_chunks(L):
            end = dp._range_end_for_stage(fake, lo)
            fake.flat_grad[written:end] += mb2[written:end]
            written = end
            dp.stages_done(fake, hi, lo)
        dp.finish()
        both = [torch.zeros_like(mb1) for _ in range(world)]
        dist.all_gather(both, mb1 + mb2)
        ok = bool(torch.allclose(fake.flat_grad, sum(both) / world, atol=1e-6))
        for p, gr in zip(probe.parameters(), fg):
            allg = [torch.zeros_like(gr) for _ in range(world)]
            dist.all_gather(allg, gr)
            ok = ok and bool(torch.allclose(p.grad, sum(allg) / world, atol=1e-6))
        # a second synchronised backward without zeroing in between is flagged
        import warnings
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            dp.stage_chunks(L)
        q.put((rank, ok, len(w) == 1 and "no_sync" in str(w[0].message)))
    finally:
        dist.destroy_process_group()


def test_no_sync_accumulation_and_foreign_parameters_gloo_world2():
    """ADVICE r1 (parallel.py:79): gradient accumulation must not re-reduce earlier micro-batches, and parameters outside
    the engine's flat buffer (extra classifiers, replaced heads) must be reduced too."""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_nosync_foreign, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok and warned for _, ok, warned in res), res


def test_torch_ddp_wrapper_is_refused():
    class FakeDDP(torch.nn.parallel.DistributedDataParallel):
        def __init__(self):  # no process group needed for the isinstance check
            torch.nn.Module.__init__(self)
    import unittest.mock as mock
    with mock.patch.object(dist, "is_initialized", return_value=True):
        with pytest.raises(RuntimeError, match="DistributedDataParallel"):
            DataParallel(FakeDDP())
