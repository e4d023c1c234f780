"""Data-parallel training over the GPUs of one box: one process per GPU, gradients all-reduced
bucket by bucket over NCCL/NVLink on a side stream WHILE the backward kernels of earlier layers
are still running.

The reference delegates this to an external trainer (omega) that wraps the model in torch DDP
(evidence: `module.`-prefixed checkpoints, examples/evaluation.py:137-138; per-rank batch =
batch_size // world_size, examples/CIFAR100.py:22).  torch DDP's hook-based overlap cannot see
inside a single fused backward, so the engine itself reports finished backward stages and this
class launches the collective for the flat-gradient range they completed.  The flat gradient
buffer is laid out in reverse execution order (engine.py), so each bucket is one contiguous slice.
Gradients are SUMMED here; the 1/world factor is folded into the fused AdamW (grad_scale).
"""
import torch
import torch.distributed as dist


class DataParallel:
    """dp = DataParallel(model, optimizer=opt, bucket_layers=2).  Use the model as usual; call
    dp.finish() (or opt.step() through dp.step()) after backward."""

    def __init__(self, model, optimizer=None, process_group=None, bucket_layers=2, average_in_optimizer=True):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised before DataParallel(model)")
        self.model = model
        self.engine = model._nrv
        self.pg = process_group
        self.world = dist.get_world_size(process_group)
        self.bucket_layers = max(1, int(bucket_layers))
        self.engine.ddp = self
        self.optimizer = optimizer
        self.broadcast = True
        self.comm_stream = None
        self.pending = []
        self.ranges = []          # (start, end) of every bucket issued (introspection / tests)
        self._done_upto = 0
        self.average_in_optimizer = average_in_optimizer and optimizer is not None and hasattr(optimizer, "grad_scale")
        if self.average_in_optimizer:
            optimizer.grad_scale = 1.0 / self.world
        self.broadcast_parameters()

    # -- parameters start identical on every rank (DDP does the same at construction)
    def broadcast_parameters(self):
        with torch.no_grad():
            for p in self.model.parameters():
                dist.broadcast(p.data, src=0, group=self.pg)
        if self.engine.flat_param is not None:
            self.engine.shadow_valid = False

    # -- called by Engine.backward
    def stage_chunks(self, L):
        """[(hi, lo)] backward stage ranges, each followed by one all-reduce."""
        chunks = []
        hi = L
        lo = max(L - self.bucket_layers, 0)
        chunks.append((hi, lo))          # head-side stage L together with the last layers
        nxt = lo - 1
        while nxt >= 0:
            lo = max(nxt - self.bucket_layers + 1, 0)
            chunks.append((nxt, lo))
            nxt = lo - 1
        last_hi, last_lo = chunks[-1]
        chunks[-1] = (last_hi, -1)       # embedding stage rides with layer 0
        return chunks

    def _range_end_for_stage(self, eng, lo):
        """End offset (exclusive) of the flat-gradient prefix that is final once stages >= lo ran."""
        if lo <= -1:
            return eng.flat_grad.numel()
        # b_fc2 of layer lo-1 is produced by layer lo's LN1 backward, everything else of layer lo-1 is not
        # final yet: the safe prefix ends where layer lo-1 starts.
        name = "l%d.w_fc2" % (lo - 1)
        if lo - 1 >= 0 and name in eng.slots:
            return eng.slots[name].offset
        return eng.flat_grad.numel()

    def head_done(self, eng):
        pass  # the head's gradients ride with the first bucket

    def stages_done(self, eng, hi, lo):
        end = self._range_end_for_stage(eng, lo)
        start = self._done_upto
        if end > start:
            seg = eng.flat_grad[start:end]
            if seg.is_cuda:
                if self.comm_stream is None:
                    self.comm_stream = torch.cuda.Stream(device=seg.device)
                cur = torch.cuda.current_stream()
                ev = torch.cuda.Event()
                ev.record(cur)
                self.comm_stream.wait_event(ev)      # the bucket's dW kernels have been enqueued before `ev`
                with torch.cuda.stream(self.comm_stream):
                    work = dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            else:
                work = dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            self.pending.append(work)
            self.ranges.append((start, end))
            self._done_upto = end
        if lo <= -1:
            self._done_upto = 0  # ready for the next backward

    def finish(self):
        """Make the current stream wait for every outstanding bucket."""
        for w in self.pending:
            w.wait()
        self.pending = []
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        if not self.average_in_optimizer and self.engine.flat_grad is not None:
            self.engine.flat_grad.mul_(1.0 / self.world)

    def step(self):
        self.finish()
        if self.optimizer is not None:
            self.optimizer.step()

    def __call__(self, *a, **k):
        return self.model(*a, **k)
