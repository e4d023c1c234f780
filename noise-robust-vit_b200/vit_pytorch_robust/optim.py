"""Fused AdamW over the engine's flat buffers (nrv_adamw): one kernel per contiguous run of
parameters instead of torch's per-tensor foreach ops; also rewrites the bf16 shadow the tensor
cores read.  Semantics = torch.optim.AdamW (examples/CIFAR100.py:90-97): decoupled weight decay,
bias correction, eps added after the sqrt.  Per-group lr / weight_decay are honoured
(examples/simpler_randomlabel.py:262-277); lr may change every step (schedulers).
"""
import torch

from . import _abi


def _runs(params):
    """Group parameters into contiguous runs of an engine's flat buffer.
    Returns ([(engine, start, end)], [foreign params])."""
    by_engine = {}
    foreign = []
    for p in params:
        slot = getattr(p, "_nrv_slot", None)
        if slot is None or slot[0].flat_param is None or \
                p.data_ptr() != slot[0].flat_param.data_ptr() + 4 * slot[1]:
            foreign.append(p)
            continue
        eng = slot[0]
        s = next(s for s in eng.slots.values() if s.param is p)
        by_engine.setdefault(id(eng), (eng, []))[1].append((s.offset, s.offset + s.padded))
    runs = []
    for eng, segs in by_engine.values():
        segs.sort()
        cur_s, cur_e = segs[0]
        for s, e in segs[1:]:
            if s == cur_e:
                cur_e = e
            else:
                runs.append((eng, cur_s, cur_e))
                cur_s, cur_e = s, e
        runs.append((eng, cur_s, cur_e))
    return runs, foreign


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2,
                 max_grad_norm=None):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.max_grad_norm = max_grad_norm
        self._step = 0
        self._engine_state = {}   # id(engine) -> (m, v) flat fp32
        self._scratch = None
        self.grad_scale = 1.0     # e.g. 1/world_size folded into the update (set by DataParallel)

    def zero_grad(self, set_to_none=False):
        """Zeroes the flat gradient buffers in one memset per engine (param.grad stay attached)."""
        seen = set()
        for g in self.param_groups:
            for p in g["params"]:
                slot = getattr(p, "_nrv_slot", None)
                if slot is not None and slot[0].flat_grad is not None and p.grad is not None:
                    if id(slot[0]) not in seen:
                        slot[0].flat_grad.zero_()
                        seen.add(id(slot[0]))
                        if slot[0].ddp is not None:
                            slot[0].ddp.on_zero_grad()
                elif p.grad is not None:
                    if set_to_none:
                        p.grad = None
                    else:
                        p.grad.zero_()

    @staticmethod
    def _layout(eng):
        return tuple((n, eng.slots[n].offset, eng.slots[n].numel) for n in eng.order)

    def _mv(self, eng):
        """Flat Adam moments of one engine (same layout as its flat_param).  When the engine rebuilt its flat buffers
        (head replaced, model moved) the moments are carried over slot by slot instead of restarting from zero."""
        st = self._engine_state.get(id(eng))
        lay = self._layout(eng)
        if st is not None and st[2] == lay and st[0].device == eng.flat_param.device:
            return st
        m, v = torch.zeros_like(eng.flat_param), torch.zeros_like(eng.flat_param)
        if st is not None:
            old = {n: (o, k) for n, o, k in st[2]}
            for n, o, k in lay:
                if n in old and old[n][1] == k:
                    oo = old[n][0]
                    m[o:o + k].copy_(st[0][oo:oo + k])
                    v[o:o + k].copy_(st[1][oo:oo + k])
        st = (m, v, lay)
        self._engine_state[id(eng)] = st
        return st

    # ---- checkpoints: interchangeable with torch.optim.AdamW (state[p] = {step, exp_avg, exp_avg_sq}) --------------
    def _engine_slot(self, p):
        slot = getattr(p, "_nrv_slot", None)
        if slot is None or slot[0].flat_param is None or p.data_ptr() != slot[0].flat_param.data_ptr() + 4 * slot[1]:
            return None
        eng = slot[0]
        return eng, next(s for s in eng.slots.values() if s.param is p)

    def state_dict(self):
        """torch.optim.AdamW's format: the moments of engine-backed parameters are exported as per-parameter exp_avg /
        exp_avg_sq tensors (copies of their slices of the flat buffers) and the shared step count as `step`."""
        sd = super().state_dict()
        idx = 0
        for group in self.param_groups:
            for p in group["params"]:
                es = self._engine_slot(p)
                if es is not None and id(es[0]) in self._engine_state and self._step > 0:
                    eng, s = es
                    m, v = self._mv(eng)[:2]
                    sd["state"][idx] = {"step": torch.tensor(float(self._step)),
                                        "exp_avg": eng._view(m, s).detach().clone(),
                                        "exp_avg_sq": eng._view(v, s).detach().clone()}
                elif idx in sd["state"]:
                    sd["state"][idx] = dict(sd["state"][idx], step=torch.tensor(float(self._step)))
                idx += 1
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)     # param_groups + per-parameter state (cast to the parameters' device)
        step = 0
        with torch.no_grad():
            for group in self.param_groups:
                for p in group["params"]:
                    st = self.state.get(p)
                    if not st:
                        continue
                    step = max(step, int(float(st.get("step", 0))))
                    es = self._engine_slot(p)
                    if es is None and getattr(p, "_nrv_slot", None) is not None and p.is_cuda:
                        p._nrv_slot[0].ensure_flat(p.device)     # model moved after construction: rebuild, then retry
                        es = self._engine_slot(p)
                    if es is not None and "exp_avg" in st:
                        eng, s = es
                        m, v = self._mv(eng)[:2]
                        eng._view(m, s).copy_(st["exp_avg"])
                        eng._view(v, s).copy_(st["exp_avg_sq"])
                        del self.state[p]                         # lives in the flat buffers from here on
        self._step = step

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _abi.load()
        self._step += 1
        stream = _abi.stream_ptr()
        coef_ptr = None
        if self.max_grad_norm is not None:
            coef_ptr = self._clip_coef(lib, stream).data_ptr()
        engines = {}
        for group in self.param_groups:
            params = [p for p in group["params"] if p.grad is not None]
            runs, foreign = _runs(params)
            b1, b2 = group["betas"]
            for eng, s, e in runs:
                m, v = self._mv(eng)[:2]
                shadow = eng.flat_shadow.data_ptr() + 2 * s if eng.compute_dtype != torch.float32 else None
                _abi.check(lib.nrv_adamw(eng.flat_param.data_ptr() + 4 * s, m.data_ptr() + 4 * s,
                                         v.data_ptr() + 4 * s, eng.flat_grad.data_ptr() + 4 * s, shadow,
                                         e - s, group["lr"], b1, b2, group["eps"], group["weight_decay"],
                                         self._step, self.grad_scale, coef_ptr, stream), "nrv_adamw")
                engines[id(eng)] = eng
            for p in foreign:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise _abi.NrvError("FusedAdamW handles contiguous fp32 CUDA parameters only")
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                _abi.check(lib.nrv_adamw(p.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                         g.data_ptr(), None, p.numel(), group["lr"], b1, b2, group["eps"],
                                         group["weight_decay"], self._step, self.grad_scale, coef_ptr, stream),
                           "nrv_adamw")
        for eng in engines.values():
            if eng.compute_dtype != torch.float32:
                # every trainable segment got a fresh shadow from the kernel; frozen segments keep theirs
                if eng.shadow_valid:
                    eng.mark_shadow_fresh()
        return loss

    def _clip_coef(self, lib, stream):
        """Global-norm clip coefficient on the device (clip_grad_norm_ semantics, grad_max_norm of
        examples/CIFAR100.py:192) — no host synchronisation."""
        dev = None
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is not None:
                    dev = p.device
                    break
            if dev is not None:
                break
        if self._scratch is None or self._scratch.device != dev:
            self._scratch = torch.zeros(2, dtype=torch.float32, device=dev)
        self._scratch.zero_()
        for group in self.param_groups:
            params = [p for p in group["params"] if p.grad is not None]
            runs, foreign = _runs(params)
            for eng, s, e in runs:
                _abi.check(lib.nrv_sumsq(eng.flat_grad.data_ptr() + 4 * s, e - s, self._scratch.data_ptr(), stream),
                           "nrv_sumsq")
            for p in foreign:
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                _abi.check(lib.nrv_sumsq(g.data_ptr(), g.numel(), self._scratch.data_ptr(), stream), "nrv_sumsq")
        _abi.check(lib.nrv_clip_coef(self._scratch.data_ptr(), float(self.max_grad_norm), float(self.grad_scale),
                                     self._scratch.data_ptr() + 4, stream), "nrv_clip_coef")
        return self._scratch[1:]


def clip_grad_norm_(parameters, max_norm):
    """Device-side torch.nn.utils.clip_grad_norm_ for engine-backed parameters: returns the total
    norm (0-dim tensor) and scales the flat gradient buffers in place by min(1, max_norm/(norm+1e-6))."""
    params = [p for p in parameters if p.grad is not None]
    if not params:
        return torch.zeros(())
    lib = _abi.load()
    stream = _abi.stream_ptr()
    scratch = torch.zeros(2, dtype=torch.float32, device=params[0].device)
    runs, foreign = _runs(params)
    for eng, s, e in runs:
        _abi.check(lib.nrv_sumsq(eng.flat_grad.data_ptr() + 4 * s, e - s, scratch.data_ptr(), stream), "nrv_sumsq")
    for p in foreign:
        g = p.grad.contiguous()
        _abi.check(lib.nrv_sumsq(g.data_ptr(), g.numel(), scratch.data_ptr(), stream), "nrv_sumsq")
    _abi.check(lib.nrv_clip_coef(scratch.data_ptr(), float(max_norm), 1.0, scratch.data_ptr() + 4, stream), "nrv_clip_coef")
    coef = scratch[1]
    for eng, s, e in runs:
        eng.flat_grad[s:e].mul_(coef)
    for p in foreign:
        p.grad.mul_(coef)
    return scratch[0].sqrt()
