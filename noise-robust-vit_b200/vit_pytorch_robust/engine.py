"""Host side of the hot path: flat parameter / gradient / bf16-shadow buffers, pointer tables for
libnrvit, activation stash + workspace management, and the torch.autograd.Function wrappers that
the nn.Module shells (simple_vit.py, vit.py) call.

PyTorch here is plumbing only: it owns device memory, streams and the autograd edge.  Every FLOP of
the encoder forward/backward is issued by nrv_vit_forward / nrv_vit_backward / nrv_gemm.

Parameter storage
  All parameters of a model live in ONE fp32 buffer (`flat_param`), laid out in reverse execution
  order (head, final LN, layer L-1 ... layer 0, embedding) so that gradient buckets complete front
  to back during backward.  nn.Parameters are views into it (state_dict / load_state_dict /
  optimisers keep working; keys and shapes are the reference's).  `flat_grad` has the same layout
  and `param.grad` are views into it: the dW kernels accumulate straight into it.  `flat_shadow`
  is the bf16 copy the tensor cores read; it is refreshed by the fused AdamW kernel, or by one cast
  kernel whenever a parameter's version counter changed (foreign optimiser, load_state_dict).
"""
import ctypes as C
import gc
import os

import torch

from . import _abi

_ALIGN = 64  # elements; keeps every tensor 128/256-byte aligned in both fp32 and bf16 buffers


def compute_dtype_from_env():
    """NRV_CHECK=fp32 selects the fp32 check mode (3xTF32 GEMMs, fp32 activations)."""
    return torch.float32 if os.environ.get("NRV_CHECK", "").lower() in ("fp32", "f32", "1") else torch.bfloat16


class ParamSlot:
    __slots__ = ("name", "param", "offset", "numel", "padded")

    def __init__(self, name, param, offset, numel, padded):
        self.name, self.param, self.offset, self.numel, self.padded = name, param, offset, numel, padded


class StashLease:
    """Ownership of one activation stash, tied to the lifetime of the autograd node that needs it."""
    __slots__ = ("engine", "key", "buf")

    def __init__(self, engine, key, buf):
        self.engine, self.key, self.buf = engine, key, buf

    def __del__(self):
        try:
            if self.buf is not None:
                self.engine._return_stash(self.key, self.buf)
        except Exception:   # interpreter shutdown
            pass


class Engine:
    """One per model instance.  `spec` describes the architecture, `param_map()` (a callable)
    returns {engine-name: nn.Parameter or None} from the live module tree."""

    def __init__(self, spec, param_map_fn):
        self.spec = dict(spec)
        self.param_map_fn = param_map_fn
        self.flat_param = None
        self.flat_grad = None
        self.flat_shadow = None
        self.slots = {}          # name -> ParamSlot
        self.order = []          # names in flat order
        self.device = None
        self.versions = None
        self.shadow_valid = False
        self.compute_dtype = compute_dtype_from_env()
        self.attn_impl = _abi.ATTN_IMPL_AUTO
        # LayerNorm folded into the QKV / FC1 GEMMs (LN_FOLDED) or stand-alone kernels (LN_SEPARATE).  Measured on B200
        # (profiles/r2_ln_fold.txt): the fold removes both LayerNorm passes over the stream (2 x 30 us per ViT-B/16 layer at
        # B=256) and costs ~48 us in the four GEMM epilogues that apply / emit the row statistics, so large-batch inference
        # is even to +3 % (ViT-H/14) and it is the inference default there; small batches pay for the three extra launches
        # (statistics of the embedding, weight fold, memset) and training pays the forward gain back when the backward
        # re-creates the normalised rows for the weight-gradient GEMMs (LayerNorm backward 73 -> 95 us): both default to
        # the stand-alone kernels.
        self.ln_mode_infer = _abi.LN_FOLDED
        self.ln_fold_min_tokens = 8192
        self.ln_mode_train = _abi.LN_SEPARATE
        self._bufs = {}          # (B, training, dtype, ...) -> workspace (transient within one call)
        self._stash_pool = {}    # same key -> [free activation stashes]; a stash in use is owned by its autograd node
        self._keep = []          # ctypes arrays that must outlive calls
        self.pos_table = None    # SimpleViT sincos table (fp32 [n, D])
        self.ddp = None          # set by parallel.DataParallel
        self.w_patch_padded = None
        # inference forwards of small batches are launch-bound (ViT-H/14 at B=1: ~230 launches, 4.5 ms): the one C call
        # nrv_vit_forward is captured into a CUDA graph per (batch, buffers) and replayed.  NRV_NO_GRAPHS=1 disables it.
        self.graph_max_tokens = 0 if os.environ.get("NRV_NO_GRAPHS") else 8192
        self._graphs = {}

    # ------------------------------------------------------------------ layout
    def _ordered_names(self, pm):
        L = self.spec["depth"]
        names = ["head_w", "head_b", "lnf_g", "lnf_b"]
        for l in reversed(range(L)):
            for f in ("w_fc2", "b_fc2", "w_fc1", "b_fc1", "ln2_g", "ln2_b", "w_out", "b_out", "w_qkv", "b_qkv",
                      "ln1_g", "ln1_b"):
                names.append("l%d.%s" % (l, f))
        names += ["pos", "cls", "w_patch", "b_patch"]
        return [n for n in names if pm.get(n) is not None]

    def _slot_numel(self, name, p):
        if name == "head_w":  # class dimension padded to 8 rows so the head GEMM sees N % 8 == 0
            C_, D = p.shape
            return ((C_ + 7) // 8 * 8) * D
        if name == "head_b":
            return (p.numel() + 7) // 8 * 8
        if name == "w_patch":  # rows padded to a multiple of 8 elements (TMA needs 16-byte row pitch)
            D = p.shape[0]
            pdim = p.numel() // D
            return D * ((pdim + 7) // 8 * 8)
        return p.numel()

    def plan_layout(self, pm=None):
        """Pure host logic: (order, {name: ParamSlot}, total elements) of the flat buffers for the
        current module tree.  Reverse execution order, every slot padded to _ALIGN elements."""
        pm = self.param_map_fn() if pm is None else pm
        order = self._ordered_names(pm)
        off = 0
        slots = {}
        for name in order:
            p = pm[name]
            n = self._slot_numel(name, p)
            padded = (n + _ALIGN - 1) // _ALIGN * _ALIGN
            slots[name] = ParamSlot(name, p, off, p.numel(), padded)
            off += padded
        return order, slots, off

    @staticmethod
    def _view(buf, s):
        """View of slot `s` inside `buf` with the parameter's shape (row-padded for w_patch)."""
        p = s.param
        if s.name == "w_patch":
            D = p.shape[0]
            pdim = p.numel() // D
            pld = (pdim + 7) // 8 * 8
            if pld != pdim:
                rows = buf[s.offset:s.offset + D * pld].view(D, pld)[:, :pdim]
                strides = [pld]
                inner = []
                acc = 1
                for d in reversed(p.shape[1:]):
                    inner.append(acc)
                    acc *= d
                return rows.as_strided(tuple(p.shape), tuple(strides + list(reversed(inner))), rows.storage_offset())
        return buf[s.offset:s.offset + s.numel].view(p.shape)

    def ensure_flat(self, device):
        """(Re)build the flat buffers if the module's parameters are not views of them (first use,
        .to()/.cuda(), load via assignment ...)."""
        pm = self.param_map_fn()
        ok = self.flat_param is not None and self.device == device
        if ok:
            base = self.flat_param.data_ptr()
            for name in self.order:
                s = self.slots[name]
                p = pm.get(name)
                if p is not s.param or p.data_ptr() != base + 4 * s.offset or p.dtype != torch.float32:
                    ok = False
                    break
            if ok and len(self.order) != len([n for n in pm if pm[n] is not None]):
                ok = False
        if ok:
            return
        if device.type != "cuda":
            raise _abi.NrvError("the vit_pytorch_robust hot path runs on an sm_100 GPU only: move the model "
                                "and the input to cuda (there is no CPU fallback)")
        _abi.init(device)
        order, slots, off = self.plan_layout(pm)
        flat = torch.zeros(off, dtype=torch.float32, device=device)
        grad = torch.zeros(off, dtype=torch.float32, device=device)
        with torch.no_grad():
            for name in order:
                s = slots[name]
                p = s.param
                view = self._view(flat, s)
                view.copy_(p.detach().to(device=device, dtype=torch.float32))
                old_grad = p.grad
                p.data = view
                gview = self._view(grad, s)
                if old_grad is not None:
                    gview.copy_(old_grad.to(device=device, dtype=torch.float32))
                    p.grad = gview
                p._nrv_slot = (self, s.offset, s.numel)
        self.flat_param, self.flat_grad = flat, grad
        self.flat_shadow = torch.zeros(off, dtype=torch.bfloat16, device=device)
        self.slots, self.order, self.device = slots, order, device
        self.shadow_valid = False
        self.versions = None
        self._bufs.clear()
        self._stash_pool.clear()
        self._graphs.clear()
        self._ptr_cache = {}
        self.pos_table = None

    # ------------------------------------------------------------------ shadows
    def _current_versions(self):
        return tuple(self.slots[n].param._version for n in self.order)

    def refresh_shadow(self, force=False):
        v = self._current_versions()
        if force or not self.shadow_valid or v != self.versions:
            lib = _abi.load()
            _abi.check(lib.nrv_cast_bf16(self.flat_param.data_ptr(), self.flat_shadow.data_ptr(),
                                         self.flat_param.numel(), _abi.stream_ptr()), "nrv_cast_bf16")
            self.versions = v
            self.shadow_valid = True
            self.w_patch_padded = None

    def mark_shadow_fresh(self):
        """Called by FusedAdamW after its kernel rewrote parameters and shadow together."""
        self.versions = self._current_versions()
        self.shadow_valid = True
        self.w_patch_padded = None

    # ------------------------------------------------------------------ gradients
    def attach_grads(self):
        """Make param.grad views of flat_grad; segments whose .grad was None start from zero."""
        missing = [n for n in self.order if self.slots[n].param.requires_grad and self.slots[n].param.grad is None]
        if not missing:
            # still make sure existing grads are OUR views (a foreign tensor would not receive the kernels' output)
            for n in self.order:
                s = self.slots[n]
                g = s.param.grad
                if g is not None and g.data_ptr() != self.flat_grad.data_ptr() + 4 * s.offset:
                    view = self._view(self.flat_grad, s)
                    view.copy_(g)
                    s.param.grad = view
            return
        n_req = sum(1 for n in self.order if self.slots[n].param.requires_grad)
        if len(missing) == n_req:
            self.flat_grad.zero_()
            if self.ddp is not None:
                self.ddp.on_zero_grad()
        for n in missing:
            s = self.slots[n]
            seg = self.flat_grad[s.offset:s.offset + s.padded]
            if len(missing) != n_req:
                seg.zero_()
            s.param.grad = self._view(self.flat_grad, s)

    # ------------------------------------------------------------------ pointer tables
    def _addr(self, buf, name):
        s = self.slots.get(name)
        if s is None:
            return None
        return buf.data_ptr() + buf.element_size() * s.offset

    def _tables(self, kind):
        """kind: 'param' (weights in compute dtype, vectors fp32) or 'grad' (all fp32, NULL where the
        parameter does not require grad)."""
        L = self.spec["depth"]
        layers = (_abi.VitLayer * L)()
        wbuf = self.flat_param if (kind == "grad" or self.compute_dtype == torch.float32) else self.flat_shadow
        vbuf = self.flat_grad if kind == "grad" else self.flat_param
        if kind == "grad":
            wbuf = self.flat_grad

        def ok(name):
            s = self.slots.get(name)
            return s is not None and (kind == "param" or s.param.requires_grad)

        for l in range(L):
            for f in _abi.LAYER_FIELDS:
                name = "l%d.%s" % (l, f)
                buf = wbuf if f.startswith("w_") else vbuf
                setattr(layers[l], f, self._addr(buf, name) if ok(name) else None)
        t = _abi.VitParams()
        t.layers = C.cast(layers, C.POINTER(_abi.VitLayer))
        t.w_patch = self._addr(wbuf, "w_patch") if ok("w_patch") else None
        t.b_patch = self._addr(vbuf, "b_patch") if ok("b_patch") else None
        t.lnf_g = self._addr(vbuf, "lnf_g") if ok("lnf_g") else None
        t.lnf_b = self._addr(vbuf, "lnf_b") if ok("lnf_b") else None
        t.cls = self._addr(vbuf, "cls") if ok("cls") else None
        if "pos" in self.slots:
            t.pos = self._addr(vbuf, "pos") if ok("pos") else None
        elif kind == "param":
            t.pos = self.pos_table.data_ptr() if self.pos_table is not None else None
        return t, layers

    # ------------------------------------------------------------------ config / buffers
    def make_config(self, B, img, training, drop=None, for_backward=None):
        sp = self.spec
        c = _abi.VitConfig()
        c.batch, c.channels = B, sp["channels"]
        c.img_h, c.img_w = sp["image_size"]
        c.patch_h, c.patch_w = sp["patch_size"]
        c.dim, c.depth, c.heads, c.dim_head, c.mlp_dim = sp["dim"], sp["depth"], sp["heads"], sp["dim_head"], sp["mlp_dim"]
        c.cls_token = 1 if sp["cls_token"] else 0
        c.pool = _abi.POOL_CLS if sp["pool"] == "cls" else _abi.POOL_MEAN
        c.patch_order = _abi.PATCH_CP1P2 if sp["patch_order"] == "cp1p2" else _abi.PATCH_P1P2C
        c.qkv_bias = 1 if sp["qkv_bias"] else 0
        c.ln_eps = sp["ln_eps"]
        c.attn_mode = _abi.ATTN_SINKHORN3 if sp.get("robust") else _abi.ATTN_SOFTMAX
        c.attn_impl = self.attn_impl
        if training if for_backward is None else for_backward:
            c.ln_mode = self.ln_mode_train
        else:
            tokens = B * ((c.img_h // c.patch_h) * (c.img_w // c.patch_w) + c.cls_token)
            c.ln_mode = self.ln_mode_infer if tokens >= self.ln_fold_min_tokens else _abi.LN_SEPARATE
        c.img_dtype = _abi._dt(img)
        c.dtype = _abi.NRV_F32 if self.compute_dtype == torch.float32 else _abi.NRV_BF16
        c.training = 1 if training else 0
        if drop is not None and training:
            c.p_drop, c.p_emb_drop, c.p_attn_drop = drop["p"], drop["p_emb"], drop["p_attn"]
            c.drop_seed = drop["seed"]
        return c

    @staticmethod
    def _buf_key(cfg):
        return (cfg.batch, cfg.training, cfg.dtype, cfg.attn_impl, cfg.p_drop > 0.0, cfg.ln_mode)

    def buffers(self, cfg):
        """Transient workspace of one nrv_vit_forward / nrv_vit_backward call (stream-ordered, so calls share it)."""
        key = self._buf_key(cfg)
        work = self._bufs.get(key)
        if work is None:
            lib = _abi.load()
            wb = lib.nrv_vit_workspace_bytes(C.byref(cfg))
            if wb == 0:
                raise _abi.NrvError("nrv_vit_workspace_bytes rejected the configuration: %s" %
                                    (lib.nrv_last_error() or b"").decode())
            work = torch.empty(wb, dtype=torch.uint8, device=self.device)
            # one entry per mode is enough: drop buffers of other batch sizes to bound memory
            for k in [k for k in self._bufs if k[1:] == key[1:]]:
                del self._bufs[k]
            for k in [k for k in self._stash_pool if k[1:] == key[1:] and k != key]:
                del self._stash_pool[k]
            self._bufs[key] = work
        return work

    def take_stash(self, cfg):
        """Activation stash for ONE grad-mode forward.  It belongs to the autograd node of that forward (StashLease, held
        by ctx) until the node dies, so a second forward before the first backward (two views, siamese / distillation
        wrappers, loss = f(a) + f(b)) gets its own buffer instead of overwriting the first one's activations."""
        key = self._buf_key(cfg)
        free = self._stash_pool.get(key)
        if free:
            return StashLease(self, key, free.pop())
        sb = _abi.load().nrv_vit_stash_bytes(C.byref(cfg))
        return StashLease(self, key, torch.empty(max(sb, 16), dtype=torch.uint8, device=self.device))

    def _return_stash(self, key, buf):
        # keep at most one idle stash per mode (the steady state of a training loop); extra ones go back to torch's allocator
        if key in self._bufs and buf.device == self.device and not self._stash_pool.get(key):
            self._stash_pool[key] = [buf]

    def ensure_pos_table(self):
        sp = self.spec
        if "pos" in self.slots or self.pos_table is not None:
            return
        h = sp["image_size"][0] // sp["patch_size"][0]
        w = sp["image_size"][1] // sp["patch_size"][1]
        if sp["dim"] % 4 != 0:
            raise AssertionError("feature dimension must be multiple of 4 for sincos emb")
        self.pos_table = torch.empty(h * w, sp["dim"], dtype=torch.float32, device=self.device)
        _abi.check(_abi.load().nrv_posemb_sincos_2d(self.pos_table.data_ptr(), h, w, sp["dim"], 10000.0,
                                                    _abi.stream_ptr()), "nrv_posemb_sincos_2d")

    # ------------------------------------------------------------------ forward / backward
    def check_input(self, img):
        sp = self.spec
        if not img.is_cuda:
            raise _abi.NrvError("input images must be CUDA tensors (no CPU fallback)")
        if img.dim() != 4 or img.shape[1] != sp["channels"]:
            raise ValueError("expected images of shape [B, %d, H, W], got %s" % (sp["channels"], tuple(img.shape)))
        if img.dtype not in (torch.float32, torch.bfloat16):
            img = img.float()
        return img.contiguous()

    def forward(self, img, training, drop=None, for_backward=None):
        """img [B,C,H,W] -> feat [B, D] in compute dtype (final-LN'ed pooled token).
        drop: None, or {"p", "p_emb", "p_attn", "seed"} for a training-mode forward with dropout.
        for_backward: whether a backward pass can follow (default: `training`); a stash kept only for introspection
        hooks keeps the inference arithmetic (LayerNorm mode)."""
        lib = _abi.load()
        self.ensure_flat(img.device)
        self.ensure_pos_table()
        if self.compute_dtype != torch.float32:
            self.refresh_shadow()
        cfg = self.make_config(img.shape[0], img, training, drop, for_backward)
        tokens = img.shape[0] * ((cfg.img_h // cfg.patch_h) * (cfg.img_w // cfg.patch_w) + cfg.cls_token)
        if not training and tokens <= self.graph_max_tokens and not torch.cuda.is_current_stream_capturing():
            return self._graphed_forward(lib, cfg, img), cfg, None
        work = self.buffers(cfg)
        lease = self.take_stash(cfg) if training else None
        ptab, keep = self._tables("param")
        feat = torch.empty(img.shape[0], self.spec["dim"], dtype=self.compute_dtype, device=img.device)
        _abi.check(lib.nrv_vit_forward(C.byref(cfg), C.byref(ptab), img.data_ptr(), feat.data_ptr(),
                                       lease.buf.data_ptr() if lease is not None else None, work.data_ptr(),
                                       _abi.stream_ptr()), "nrv_vit_forward")
        del keep
        return feat, cfg, lease

    def _graphed_forward(self, lib, cfg, img):
        """Inference forward through a captured CUDA graph (nrv_vit_forward allocates nothing and never synchronises).
        Everything the kernels address is static: the flat parameter / shadow buffers and, owned by the graph entry, a
        workspace, an input and an output buffer.  Parameter updates happen in place, so an entry stays valid until the
        flat buffers are rebuilt (ensure_flat clears the cache)."""
        wsrc = self.flat_param if self.compute_dtype == torch.float32 else self.flat_shadow
        key = (cfg.batch, cfg.dtype, img.dtype, cfg.attn_impl, cfg.attn_mode, wsrc.data_ptr(), self.flat_param.data_ptr(),
               self.pos_table.data_ptr() if self.pos_table is not None else 0)
        ent = self._graphs.get(key)
        if ent is None:
            wb = lib.nrv_vit_workspace_bytes(C.byref(cfg))
            if wb == 0:
                raise _abi.NrvError("nrv_vit_workspace_bytes rejected the configuration: %s" %
                                    (lib.nrv_last_error() or b"").decode())
            work = torch.empty(wb, dtype=torch.uint8, device=self.device)
            ptab, keep = self._tables("param")
            s_img = torch.empty_like(img)
            s_feat = torch.empty(img.shape[0], self.spec["dim"], dtype=self.compute_dtype, device=img.device)
            s_img.copy_(img)

            def call():
                _abi.check(lib.nrv_vit_forward(C.byref(cfg), C.byref(ptab), s_img.data_ptr(), s_feat.data_ptr(), None,
                                               work.data_ptr(), _abi.stream_ptr()), "nrv_vit_forward")
            call()                      # first-use work (function attributes, descriptor caches) stays outside the capture
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            # A cyclic-garbage collection DURING the capture may run destructors that call into CUDA (graphs, events and
            # buffers of models that died earlier) and invalidates it (cudaErrorStreamCaptureInvalidated; seen once the test
            # suite had left enough garbage).  torch.cuda.graph no longer collects on entry: collect now, then keep the
            # collector off until the capture has ended.
            gc.collect()
            gc_was_on = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    call()
            finally:
                if gc_was_on:
                    gc.enable()
            if len(self._graphs) >= 8:  # bound the private buffers kept alive
                self._graphs.pop(next(iter(self._graphs)))
            ent = (g, s_img, s_feat, work, keep)
            self._graphs[key] = ent
        g, s_img, s_feat = ent[0], ent[1], ent[2]
        s_img.copy_(img)
        g.replay()
        return s_feat.clone()

    def backward(self, cfg, img, dfeat, lease):
        """`lease` is the StashLease the forward of this very autograd node filled (never re-allocated here)."""
        lib = _abi.load()
        if lease is None or lease.buf is None:
            raise _abi.NrvError("backward without the activation stash of its forward pass")
        stash = lease.buf
        work = self.buffers(cfg)
        self.attach_grads()
        ptab, k1 = self._tables("param")
        gtab, k2 = self._tables("grad")
        L = self.spec["depth"]
        stages = [(L, -1)] if self.ddp is None else self.ddp.stage_chunks(L)
        for hi, lo in stages:
            # data parallel: the library records the bucket's start event in front of the LayerNorm backward that closes the
            # chunk (nrv_vit_backward_marker), so the all-reduce begins under that kernel instead of beside the next GEMM
            marker = self.ddp.marker_event(cfg) if self.ddp is not None and hasattr(self.ddp, "marker_event") else None
            if marker is not None:
                _abi.check(lib.nrv_vit_backward_marker(marker.cuda_event), "nrv_vit_backward_marker")
            _abi.check(lib.nrv_vit_backward(C.byref(cfg), C.byref(ptab), C.byref(gtab), img.data_ptr(),
                                            dfeat.data_ptr(), stash.data_ptr(), work.data_ptr(), hi, lo,
                                            _abi.stream_ptr()), "nrv_vit_backward")
            if self.ddp is not None:
                if marker is not None:
                    self.ddp.stages_done(self, hi, lo, marker=marker, early=cfg.ln_mode != _abi.LN_FOLDED or cfg.p_drop > 0)
                else:
                    self.ddp.stages_done(self, hi, lo)
        del k1, k2

    # head -------------------------------------------------------------------------------------
    def head_forward(self, feat):
        """logits fp32 [B, C] = feat @ head_w^T + head_b  (simple_vit.py:136 Linear ; vit.py:265)."""
        s = self.slots["head_w"]
        Cn, D = s.param.shape
        Cp = (Cn + 7) // 8 * 8
        wbuf = self.flat_param if self.compute_dtype == torch.float32 else self.flat_shadow
        w = wbuf[s.offset:s.offset + Cp * D].view(Cp, D)
        bias = None
        if "head_b" in self.slots:
            sb = self.slots["head_b"]
            bias = self.flat_param[sb.offset:sb.offset + Cp]
        logits = torch.empty(feat.shape[0], Cp, dtype=torch.float32, device=feat.device)
        _abi.gemm(feat, w, logits, bias=bias)
        return logits if Cp == Cn else logits[:, :Cn]

    def head_backward(self, feat, dlogits):
        """dlogits fp32 [B, C] -> dfeat (compute dtype) ; accumulates head_w / head_b gradients."""
        lib = _abi.load()
        self.attach_grads()
        s = self.slots["head_w"]
        Cn, D = s.param.shape
        Cp = (Cn + 7) // 8 * 8
        B = feat.shape[0]
        dl = dlogits
        if Cp != Cn or not dl.is_contiguous():
            pad = torch.zeros(B, Cp, dtype=torch.float32, device=feat.device)
            pad[:, :Cn] = dl
            dl = pad
        if self.compute_dtype == torch.float32:
            dlc = dl
        else:
            dlc = torch.empty(B, Cp, dtype=torch.bfloat16, device=feat.device)
            _abi.check(lib.nrv_cast_bf16(dl.data_ptr(), dlc.data_ptr(), B * Cp, _abi.stream_ptr()), "nrv_cast_bf16")
        wbuf = self.flat_param if self.compute_dtype == torch.float32 else self.flat_shadow
        w = wbuf[s.offset:s.offset + Cp * D].view(Cp, D)
        dfeat = torch.empty_like(feat)
        _abi.gemm(dlc, w, dfeat, b_layout=_abi.NRV_MN_MAJOR, M=B, N=D, K=Cp)
        if s.param.requires_grad:
            gw = self.flat_grad[s.offset:s.offset + Cp * D].view(Cp, D)
            _abi.gemm(dlc, feat, gw, a_layout=_abi.NRV_MN_MAJOR, b_layout=_abi.NRV_MN_MAJOR, epi=_abi.EPI_ATOMIC_F32,
                      M=Cp, N=D, K=B)
        if "head_b" in self.slots and self.slots["head_b"].param.requires_grad:
            sb = self.slots["head_b"]
            gb = self.flat_grad[sb.offset:sb.offset + Cp]
            nb = lib.nrv_colsum_workspace(B, Cp)
            ws = torch.empty(nb, dtype=torch.uint8, device=feat.device)
            _abi.check(lib.nrv_colsum(dlc.data_ptr(), Cp, B, Cp, _abi._dt(dlc), gb.data_ptr(), ws.data_ptr(), nb,
                                      _abi.stream_ptr()), "nrv_colsum")
        if self.ddp is not None:
            self.ddp.head_done(self)
        return dfeat


class EncoderFn(torch.autograd.Function):
    """img -> feat.  Parameters are passed so autograd knows the dependency; their gradients are
    accumulated into param.grad by the kernels (views of Engine.flat_grad), not returned."""

    @staticmethod
    def forward(ctx, engine, img, drop, want_grad, keep_stash, *params):
        # want_grad is decided by the caller: inside Function.forward grad mode is always off, and
        # ctx.needs_input_grad is True for every parameter that requires grad even under torch.no_grad()
        # a train()-mode forward with dropout draws masks even under no_grad, as nn.Dropout does;
        # keep_stash: forward hooks want per-layer tensors, which only the training-layout stash holds
        feat, cfg, lease = engine.forward(img, training=want_grad or drop is not None or keep_stash, drop=drop,
                                          for_backward=want_grad or drop is not None)
        ctx.engine, ctx.cfg, ctx.img, ctx.lease = engine, cfg, img, lease
        ctx.img_version = img._version    # backward re-reads the image (patch-embedding weight gradient gathers it by TMA)
        if keep_stash:
            engine._introspect = (cfg, lease)
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        eng = ctx.engine
        if ctx.img._version != ctx.img_version:
            raise RuntimeError("the input images were modified in place between forward and backward (version %d -> %d); "
                               "the patch-embedding weight gradient reads them" % (ctx.img_version, ctx.img._version))
        if dfeat.dtype != eng.compute_dtype:
            dfeat = dfeat.to(eng.compute_dtype)
        eng.backward(ctx.cfg, ctx.img, dfeat.contiguous(), ctx.lease)
        return (None, None, None, None, None) + (None,) * (len(ctx.needs_input_grad) - 5)


class HeadFn(torch.autograd.Function):
    """feat -> logits (fp32) through nrv_gemm; head gradients accumulate into the flat buffer."""

    @staticmethod
    def forward(ctx, engine, feat, *params):
        ctx.engine = engine
        ctx.save_for_backward(feat)
        return engine.head_forward(feat)

    @staticmethod
    def backward(ctx, dlogits):
        (feat,) = ctx.saved_tensors
        dfeat = ctx.engine.head_backward(feat, dlogits.float())
        return (None, dfeat) + (None,) * (len(ctx.needs_input_grad) - 2)


def dropout_request(module_training, p=0.0, p_emb=0.0, p_attn=0.0, robust=False):
    """None unless this is a train()-mode forward with some dropout probability > 0.  The seed comes from
    torch's default CPU generator (so torch.manual_seed makes runs reproducible) without a device sync; the
    element masks are a pure function of (seed, layer, site, index): see nrv_dropout in include/nrvit.h.
    p_attn > 0 (dropout on the attention probabilities): the general tcgen05 attention kernels draw the mask in bf16
    (dh <= 80, <= 384 tokens), the CUDA-core kernels in the fp32 check mode and beyond."""
    if not module_training or max(p, p_emb, p_attn) <= 0.0:
        return None
    if p_attn > 0.0 and robust:
        raise NotImplementedError(
            "attention dropout (p=%g) together with robust=True (Sinkhorn attention) is not implemented for training "
            "(eval() works; there is no unfused fallback)" % p_attn)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    return {"p": float(p), "p_emb": float(p_emb), "p_attn": float(p_attn), "seed": seed}


class StashView:
    """Read access to the per-layer tensors of one forward pass for the introspection wrappers of the reference
    (recorder.py:28-31 hooks `Attention.attend`, extractor.py:50-59 hooks `vit.transformer`).  The fused encoder never
    calls its parameter-holder sub-modules, so the model shells use this view to hand the hooked modules the tensors
    the reference modules would have produced.  Everything returned is a fresh tensor (the stash is recycled)."""

    def __init__(self, engine, cfg, lease):
        self.engine, self.cfg, self.lease = engine, cfg, lease
        sp = engine.spec
        self.B = cfg.batch
        self.N = (cfg.img_h // cfg.patch_h) * (cfg.img_w // cfg.patch_w) + cfg.cls_token
        self.H, self.dh, self.D, self.L = sp["heads"], sp["dim_head"], sp["dim"], sp["depth"]

    def _tensor(self, what, index, shape):
        off, nbytes = C.c_size_t(), C.c_size_t()
        _abi.check(_abi.load().nrv_vit_stash_tensor(C.byref(self.cfg), what, index, C.byref(off), C.byref(nbytes)),
                   "nrv_vit_stash_tensor")
        raw = self.lease.buf[off.value:off.value + nbytes.value]
        return raw.view(self.engine.compute_dtype).view(shape)

    def stream(self, k):
        """Residual stream [B, N, D] (fp32 copy): k = 2l enters layer l, k = 2l + 1 sits between its two branches,
        k = 2 * depth is the transformer's output."""
        return self._tensor(_abi.STASH_STREAM, k, (self.B, self.N, self.D)).float()

    def attention_probs(self, layer):
        """fp32 [B, H, N, N]: what the reference's `attend` returns in this layer (softmax, or softmax + Sinkhorn)."""
        lib = _abi.load()
        qkv = self._tensor(_abi.STASH_QKV, layer, (self.B, self.N, 3 * self.H * self.dh))
        B, N, H, dh = self.B, self.N, self.H, self.dh
        dev = qkv.device
        probs = torch.empty(B, H, N, N, dtype=torch.float32, device=dev)
        stats = torch.empty(lib.nrv_attn_stats_elems(B, N, H, _abi.ATTN_SINKHORN3), dtype=torch.float32, device=dev)
        nb = lib.nrv_attn_fwd_workspace(B, N, H, dh, _abi.ATTN_SINKHORN3)
        ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
        _abi.check(lib.nrv_attn_probs(qkv.data_ptr(), probs.data_ptr(), stats.data_ptr(), B, N, H, dh, float(dh) ** -0.5,
                                      self.cfg.attn_mode, _abi._dt(qkv), ws.data_ptr(), nb, _abi.stream_ptr()), "nrv_attn_probs")
        return probs


def has_forward_hooks(module):
    """True if calling `module` would run a user hook (module-level or global)."""
    from torch.nn.modules import module as _m
    return bool(module._forward_hooks or module._forward_pre_hooks or _m._global_forward_hooks or
                _m._global_forward_pre_hooks)


def emit(module, inp, out):
    """Run `module.__call__` so that every registered hook fires with (input, output) = (inp, out) while the module's own
    arithmetic stays inside the fused encoder: the parameter-holder shells return `_nrv_out` when it is set."""
    module._nrv_out = out
    try:
        res = module(inp)
    finally:
        module._nrv_out = None
    if res is not out:
        raise NotImplementedError(
            "a forward hook on %s replaced the module output; the fused encoder cannot feed a modified activation back "
            "into the pass (hooks may observe, not rewrite)" % type(module).__name__)


def run_model(engine, img, with_head=True, drop=None, introspect=None):
    """Shared forward of both model families.  introspect: None, or a callable(StashView) the model shell passes when forward
    hooks are registered on its sub-modules; it runs right after the encoder pass."""
    if img.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError(
            "the fused encoder does not produce the gradient with respect to the input images (the patch-embedding dX is "
            "not computed; reference: autograd through simple_vit.py:126-131 / vit.py:323-331): detach() the input, or "
            "use the reference module for saliency / adversarial evaluations")
    img = engine.check_input(img)
    engine.ensure_flat(img.device)
    engine.last_dropout = drop
    enc_params = [engine.slots[n].param for n in engine.order if not n.startswith("head_")]
    # the activation stash (and the GELU' epilogue) are only paid for when a backward pass can follow
    want_grad = torch.is_grad_enabled() and any(p.requires_grad for p in enc_params)
    feat = EncoderFn.apply(engine, img, drop, want_grad, introspect is not None, *enc_params)
    if introspect is not None:
        cfg, lease = engine._introspect
        engine._introspect = None
        with torch.no_grad():
            introspect(StashView(engine, cfg, lease))
        del lease
    if not with_head:
        return feat
    head_params = [engine.slots[n].param for n in engine.order if n.startswith("head_")]
    return HeadFn.apply(engine, feat, *head_params)
