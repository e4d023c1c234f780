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

The collective itself is libnrvit's (nrv_comm_*, csrc/comm.cu): a NCCL communicator created with a CTA cap -- the
backward kernels are persistent and fill every SM, so an all-reduce that asks for many CTAs only queues behind them --
on which the flat gradient buffer is registered once.  torch.distributed provides the rendezvous (the unique id is
broadcast through the existing process group) and carries the few parameters outside the flat buffer; CPU tensors
(the gloo tests of the bucket logic) go through torch.distributed.all_reduce.
"""
import contextlib
import ctypes as C
import os
import warnings

import torch
import torch.distributed as dist

from . import _abi


class DataParallel:
    """dp = DataParallel(model, optimizer=opt, bucket_layers=2).  Use the model as usual; call
    dp.finish() (or opt.step() through dp.step()) after backward."""

    def __init__(self, model, optimizer=None, process_group=None, bucket_layers=2, average_in_optimizer=True,
                 extra_modules=(), comm="nrv", max_ctas=0, grad_dtype=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised before DataParallel(model)")
        if isinstance(model, torch.nn.parallel.DistributedDataParallel):
            # torch DDP's hook-based overlap cannot see inside the single fused backward: it would reduce
            # everything after the whole backward, and a second time here
            raise RuntimeError("pass the bare vit_pytorch_robust model, not a torch DistributedDataParallel wrapper: "
                               "DataParallel issues the bucketed all-reduce itself from inside the fused backward")
        self.model = model
        # modules trained next to the encoder (examples/simpler_randomlabel.py:183-220: classifier / extra_classifier
        # on top of heads.head = Identity): their parameters are broadcast and their gradients reduced in finish()
        self.extra_modules = list(extra_modules)
        self.sync = True              # False inside no_sync(): gradient accumulation without a collective
        self._reduced_since_zero = False
        self._warned_accum = False
        self.engine = model._nrv
        self.pg = process_group
        self.world = dist.get_world_size(process_group)
        self.bucket_layers = max(1, int(bucket_layers))
        self.engine.ddp = self
        self.optimizer = optimizer
        self.broadcast = True
        self.comm_stream = None
        # "nrv": nrv_comm_allreduce_bucket on a communicator of our own (CTA cap, registered buffer); "torch": the process
        # group's all_reduce.  NRV_COMM=torch switches for A/B runs.
        self.comm_kind = os.environ.get("NRV_COMM", comm)
        # CTA cap of the communicator, 0 = NCCL's choice.  Measured on 2 x B200 (profiles/r2_ddp_comm.txt): a cap makes the
        # collective longer and the step slower (4 CTAs -3.5 %, 8 CTAs -1.5 %), so the default leaves it alone.
        self.max_ctas = int(os.environ.get("NRV_COMM_MAX_CTAS", max_ctas))
        # "bf16": a bucket is cast to bf16, all-reduced and cast back on the side stream (half the bytes on the wire, the
        # NCCL kernel holds its SMs half as long); "fp32": the flat buffer itself.  Default: the engine's compute dtype,
        # i.e. fp32 exchange in the fp32 check mode.
        self.grad_dtype = os.environ.get("NRV_COMM_GRAD_DTYPE", grad_dtype)
        # NRV_DDP_EARLY=1: start a bucket's all-reduce at the library's marker (in front of the closing LayerNorm backward)
        # instead of at the end of the chunk.  Off by default: no gain at 2 GPUs (profiles/r2d_ddp_early_start.txt)
        self.early_start = os.environ.get("NRV_DDP_EARLY", "0") == "1"
        self._bf16_scratch = None
        self._comm = None            # nrv_comm*
        self._comm_reg = None        # (registration handle, data_ptr of the registered flat_grad)
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
            for p in self._all_parameters():
                dist.broadcast(p.data, src=0, group=self.pg)
        if self.engine.flat_param is not None:
            self.engine.shadow_valid = False

    def _all_parameters(self):
        seen = set()
        for mod in [self.model] + self.extra_modules:
            for p in mod.parameters():
                if id(p) not in seen:
                    seen.add(id(p))
                    yield p

    def _foreign_parameters(self):
        """requires_grad parameters that do not live in the engine's flat buffer: replaced / representation_size heads
        (vit._fusable_head() False) and the extra modules."""
        eng = self.engine
        base = eng.flat_param.data_ptr() if eng.flat_param is not None else None
        end = base + 4 * eng.flat_param.numel() if base is not None else None
        for p in self._all_parameters():
            inside = base is not None and p.is_cuda and base <= p.data_ptr() < end
            if p.requires_grad and not inside:
                yield p

    @contextlib.contextmanager
    def no_sync(self):
        """Gradient accumulation: backward passes inside this context only accumulate into the local flat gradient
        buffer; the first backward outside it all-reduces the accumulated sum (torch DDP's no_sync contract)."""
        prev, self.sync = self.sync, False
        try:
            yield self
        finally:
            self.sync = prev

    def on_zero_grad(self):
        """Called by FusedAdamW.zero_grad / Engine.attach_grads when the flat gradient buffer restarts from zero."""
        self._reduced_since_zero = False

    # -- the library's communicator
    def _ensure_comm(self, device):
        if self._comm is not None:
            return self._comm
        lib = _abi.init(device)
        nbytes = lib.nrv_comm_unique_id_bytes()
        rank = dist.get_rank(self.pg)
        uid = (C.c_ubyte * nbytes)()
        if rank == 0:
            _abi.check(lib.nrv_comm_get_unique_id(uid, nbytes), "nrv_comm_get_unique_id")
        t = torch.tensor(list(uid), dtype=torch.uint8, device=device)
        dist.broadcast(t, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)
        raw = bytes(t.cpu().tolist())
        handle = C.c_void_p()
        _abi.check(lib.nrv_comm_init(raw, nbytes, self.world, rank, self.max_ctas, C.byref(handle)), "nrv_comm_init")
        self._comm = handle
        return handle

    def _register(self, eng):
        """The flat gradient buffer is registered once with the communicator (re-registered if the engine rebuilt it)."""
        ptr = eng.flat_grad.data_ptr()
        if self._comm_reg is not None and self._comm_reg[1] == ptr:
            return
        lib = _abi.load()
        if self._comm_reg is not None:
            lib.nrv_comm_deregister(self._comm, self._comm_reg[0])
        h = C.c_void_p()
        rc = lib.nrv_comm_register(self._comm, ptr, eng.flat_grad.numel() * 4, C.byref(h))
        self._comm_reg = (h, ptr) if rc == 0 else (None, ptr)     # registration is an optimisation: a refusal is not fatal

    def close(self):
        if self._comm is not None:
            lib = _abi.load()
            if self._comm_reg is not None and self._comm_reg[0]:
                lib.nrv_comm_deregister(self._comm, self._comm_reg[0])
            lib.nrv_comm_destroy(self._comm)
            self._comm, self._comm_reg = None, None

    def __del__(self):
        try:
            self.close()
        except Exception:   # interpreter shutdown
            pass

    # -- called by Engine.backward
    def stage_chunks(self, L):
        """[(hi, lo)] backward stage ranges, each followed by one all-reduce."""
        if not self.sync:
            return [(L, -1)]
        if self._reduced_since_zero and not self._warned_accum:
            self._warned_accum = True
            warnings.warn("DataParallel: a second synchronised backward before the gradients were zeroed all-reduces the "
                          "already reduced gradients of the earlier micro-batch again; wrap all but the last "
                          "micro-batch in dp.no_sync()")
        chunks = []
        hi = L
        lo = max(L - self.bucket_layers, 0)
        chunks.append((hi, lo))          # head-side stage L together with the last layers
        nxt = lo - 1
        while nxt >= 0:
            lo = max(nxt - self.bucket_layers + 1, 0)
            chunks.append((nxt, lo))
            nxt = lo - 1
        # The LAST bucket has nothing left to hide under: keep it to layer 0 + the embedding by giving
        # the layers above their own all-reduce
        last_hi, last_lo = chunks[-1]
        if last_hi > 0:
            chunks[-1] = (last_hi, 1)
            chunks.append((0, -1))
        else:
            chunks[-1] = (last_hi, -1)   # embedding stage rides with layer 0
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

    def marker_event(self, cfg):
        """A CUDA event for the library to record where the bucket of the next backward chunk becomes final (see
        nrv_vit_backward_marker); None when the early start is off (default; NRV_DDP_EARLY=1 enables it), outside synchronised backward
        passes, or for CPU tensors (the gloo tests)."""
        if not self.sync or not self.early_start or self.engine.flat_grad is None or not self.engine.flat_grad.is_cuda:
            return None
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())     # materialises the cudaEvent_t; the library re-records it
        return ev

    def head_done(self, eng):
        pass  # the head's gradients ride with the first bucket

    def stages_done(self, eng, hi, lo, marker=None, early=False):
        """marker: event recorded by the library inside the chunk's backward call.  early=True: it sits in front of the
        LayerNorm backward of layer `lo`, so layer lo's ln1 gamma / beta are not final yet and ride with the next bucket."""
        if not self.sync:
            return
        end = self._range_end_for_stage(eng, lo)
        if marker is not None and early and lo >= 0 and ("l%d.ln1_g" % lo) in eng.slots:
            end = eng.slots["l%d.ln1_g" % lo].offset
        start = self._done_upto
        if end > start:
            seg = eng.flat_grad[start:end]
            if seg.is_cuda:
                if self.comm_stream is None:
                    self.comm_stream = torch.cuda.Stream(device=seg.device)
                if marker is not None:
                    ev = marker                      # recorded by the library where the bucket's gradients are final
                else:
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream())
                self.comm_stream.wait_event(ev)      # the bucket's dW kernels have been enqueued before `ev`
                if self.comm_kind == "nrv":
                    comm = self._ensure_comm(seg.device)
                    lib = _abi.load()
                    sp = _abi.stream_ptr(self.comm_stream)
                    gd = self.grad_dtype or ("bf16" if eng.compute_dtype == torch.bfloat16 else "fp32")
                    if gd == "bf16":
                        if self._bf16_scratch is None or self._bf16_scratch.numel() != eng.flat_grad.numel():
                            self._bf16_scratch = torch.empty(eng.flat_grad.numel(), dtype=torch.bfloat16, device=seg.device)
                        half = self._bf16_scratch[start:end]
                        _abi.check(lib.nrv_cast_bf16(seg.data_ptr(), half.data_ptr(), seg.numel(), sp), "nrv_cast_bf16")
                        _abi.check(lib.nrv_comm_allreduce_bucket(comm, half.data_ptr(), half.numel(), _abi.NRV_BF16, sp),
                                   "nrv_comm_allreduce_bucket")
                        _abi.check(lib.nrv_cast_f32(half.data_ptr(), seg.data_ptr(), seg.numel(), sp), "nrv_cast_f32")
                    else:
                        self._register(eng)
                        _abi.check(lib.nrv_comm_allreduce_bucket(comm, seg.data_ptr(), seg.numel(), _abi.NRV_F32, sp),
                                   "nrv_comm_allreduce_bucket")
                    work = None
                else:
                    with torch.cuda.stream(self.comm_stream):
                        work = dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            else:
                work = dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            if work is not None:
                self.pending.append(work)
            self.ranges.append((start, end))
            self._done_upto = end
        if lo <= -1:
            self._done_upto = 0  # ready for the next backward
            self._reduced_since_zero = True

    def finish(self):
        """Make the current stream wait for every outstanding bucket; also reduces the gradients of parameters that are
        not engine-backed (replaced heads, extra modules), which no bucket covers.  Inside no_sync() it does nothing."""
        if not self.sync:
            return
        for p in self._foreign_parameters():
            if p.grad is not None:
                dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.pg)
                if not self.average_in_optimizer:
                    p.grad.mul_(1.0 / self.world)
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
