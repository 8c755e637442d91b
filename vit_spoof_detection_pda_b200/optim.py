"""Fused Adam / AdamW + fused global-norm clipping over the module's flat parameter / gradient buffers.

Reference call sites (/root/reference/train_advanced.py): ``torch.optim.AdamW(lr=3e-4, weight_decay=0.05,
betas=(0.9, 0.999))`` :592-597; ``clip_grad_norm_(model.parameters(), 1.0)`` :334; ``optimizer.step()``
via ``scaler.step`` :335; ``optimizer.zero_grad(set_to_none=True)`` :337; README.md:140-147 names Adam with
weight_decay 1e-4 (L2), which is ``adamw=False`` here.
"""
from __future__ import annotations

import weakref

import torch

from . import _lib as L


def _owner_of(params, materialize: bool = True):
    """The vitk module the parameters belong to.  With materialize=False only the first element of the iterable is
    consumed: ``clip_grad_norm_(model.parameters(), 1.0)`` hands in a generator whose full walk costs ~0.25 ms of host
    time per step, and the flat buffers already cover every parameter of the owner."""
    if materialize:
        params = list(params)
        first = params[0] if params else None
    else:
        first = next(iter(params), None)
        params = None
    if first is None:
        raise ValueError("no parameters")
    ref = getattr(first, "_vitk_owner", None)
    owner = ref() if ref is not None else None
    if owner is None:
        raise RuntimeError("parameters do not belong to a vitk ViTFaceAntiSpoofing module")
    owner._ensure_flat()   # raises with a clear message when the model is not on CUDA yet
    return owner, params


def _gather_flat_grads(owner, params):
    """Return the flat gradient buffer; zero-copy when .grad tensors are the views backward produced."""
    g = owner.flat_grads()
    # clip_grad_norm_ followed by optimizer.step() (the reference loop, train_advanced.py:334-335) walks the 156
    # parameters once per backward, not twice: ~0.6 ms of host time per step otherwise
    serial = getattr(owner, "_bwd_serial", None)
    if serial is not None and getattr(owner, "_gathered_serial", None) == serial:
        return g
    owner._gathered_serial = serial
    base = g.data_ptr()
    foreign = []
    for p, off, n in zip(owner._param_list(), owner._offsets, owner._sizes):
        if p.grad is None:
            if p.requires_grad:
                g[off:off + n].zero_()
            continue
        if p.grad.data_ptr() != base + 4 * off:
            g[off:off + n].copy_(p.grad.reshape(-1))   # foreign gradient tensor (cloned / accumulated elsewhere)
            foreign.append((p, off, n))
    owner._foreign_grads = foreign
    return g


def _scale_grads_now(owner, g, mult: float, sumsq, max_norm: float):
    """Eager in-place g *= mult * clipcoef over the flat buffer (and back into any gradient tensor that is not a view of
    it), for callers that pair ONE fused piece with stock torch ones -- e.g. FusedGradScaler.unscale_ followed by
    torch.nn.utils.clip_grad_norm_, or this module's clip_grad_norm_ in front of a stock torch optimizer."""
    L.call("vitk_grad_scale", L.ptr(g), g.numel(), float(mult), L.ptr(sumsq), float(max_norm), L.stream_ptr())
    for p, off, n in getattr(owner, "_foreign_grads", []) or []:
        p.grad.copy_(g[off:off + n].view(p.grad.shape))


def _sumsq(owner, g):
    lib = L.load()
    if getattr(owner, "_sumsq_scratch", None) is None:
        owner._sumsq_scratch = torch.empty(lib.vitk_grad_sumsq_scratch_floats(), dtype=torch.float32, device=g.device)
        owner._sumsq = torch.zeros(1, dtype=torch.float32, device=g.device)
    with L.nvtx_range("vitk/grad_norm"):
        L.call("vitk_grad_sumsq", L.ptr(g), g.numel(), L.ptr(owner._sumsq_scratch), L.ptr(owner._sumsq), L.stream_ptr())
    return owner._sumsq


def clip_grad_norm_(parameters, max_norm: float):
    """Drop-in for ``torch.nn.utils.clip_grad_norm_`` on a vitk model: one deterministic sum-of-squares
    reduction over the flat gradient buffer.  When a FusedAdam drives the model the scaling itself is folded into
    its next step() (no extra pass over the gradients); with any other optimizer the gradients are scaled in place
    right here, like torch's.  Returns the total norm (0-dim tensor)."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    owner, params = _owner_of(parameters, materialize=False)
    g = _gather_flat_grads(owner, params)
    nvls = getattr(owner, "_nvls", None)
    if nvls is not None:
        # data parallel over NVSwitch multicast: the norm is that of the REDUCED gradient.  The reduce-scatter + partial sums
        # of squares run here; the clip coefficient is applied by the fused Adam + all-gather kernel of FusedAdam.step()
        fused_ref = getattr(owner, "_fused_opt", None)
        if fused_ref is None or fused_ref() is None:
            raise RuntimeError("DataParallel(mode='nvls') needs FusedAdam: the clip is applied inside its fused step")
        nvls.reduce()
        owner._pending_clip = ("nvls", float(max_norm))
        return nvls.total_norm(getattr(owner, "_grad_mult", 1.0))
    sumsq = _sumsq(owner, g)
    mult = getattr(owner, "_grad_mult", 1.0)     # 1/scale after a LAZY FusedGradScaler.unscale_ (folded into the Adam pass)
    norm = sumsq.sqrt().reshape(())
    norm = norm * mult if mult != 1.0 else norm
    fused_ref = getattr(owner, "_fused_opt", None)
    if fused_ref is not None and fused_ref() is not None:
        owner._pending_clip = (sumsq, float(max_norm))
    else:
        _scale_grads_now(owner, g, mult, sumsq, float(max_norm))
        owner._grad_mult = 1.0
        owner._pending_clip = None
    return norm


class FusedAdam(torch.optim.Optimizer):
    """One kernel over the flat buffers: (optional) clip scale -> Adam/AdamW update -> bf16 shadow write."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, adamw=True,
                 max_grad_norm=None, grad_mult: float = 1.0, capturable: bool = False):
        params = list(params)
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, adamw=adamw)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam supports a single param group (the reference uses one)")
        self._owner, _ = _owner_of(self.param_groups[0]["params"])
        self._owner._fused_opt = weakref.ref(self)   # clip_grad_norm_ may leave its scaling to this optimizer's pass
        self.max_grad_norm = max_grad_norm
        self.grad_mult = grad_mult
        self._step = 0
        self._m = None
        self._v = None
        # capturable: step count and learning rate live in device memory (bias corrections computed on the device), so that the
        # whole training step can be captured into a CUDA graph and replayed (GraphedTrainStep)
        self.capturable = bool(capturable)
        self._step_dev = self._lr_dev = self._hyper = self._lr_host = None

    def _ensure_state(self, flat):
        if self.capturable and self._step_dev is None:      # device-side step count / learning rate (before the step is counted)
            dev = flat.device
            self._lr_host = torch.full((1,), float("nan"), dtype=torch.float32).pin_memory()
            self._lr_dev = torch.empty(1, dtype=torch.float32, device=dev)
            self._step_dev = torch.full((1,), self._step, dtype=torch.int32, device=dev)
            self._hyper = torch.zeros(4, dtype=torch.float32, device=dev)
        nvls = getattr(self._owner, "_nvls", None)
        if nvls is not None:       # NVLS data parallel: moments exist for this rank's slices only (one pair per domain)
            if not isinstance(self._m, list):
                self._m, self._v = nvls.alloc_moments()
            return
        n = flat.numel()
        if self._m is None or isinstance(self._m, list) or self._m.device != flat.device or self._m.numel() != n:
            self._m = torch.zeros(n, dtype=torch.float32, device=flat.device)
            self._v = torch.zeros(n, dtype=torch.float32, device=flat.device)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        owner = self._owner
        grp = self.param_groups[0]
        owner._ensure_flat()
        flat = owner._flat          # (not flat_params(): that accessor waits for the NVLS background master copy)
        self._ensure_state(flat)
        g = _gather_flat_grads(owner, grp["params"])
        nvls = getattr(owner, "_nvls", None)
        if nvls is not None:
            return self._step_nvls(nvls, owner, grp, loss)
        sumsq, max_norm = None, 0.0
        if owner._pending_clip is not None:
            sumsq, max_norm = owner._pending_clip
            owner._pending_clip = None
        elif self.max_grad_norm is not None:
            clip_grad_norm_(grp["params"], self.max_grad_norm)
            sumsq, max_norm = owner._pending_clip
            owner._pending_clip = None
        # frozen parameters must not move: zero their gradient AND skip decay by masking ranges
        frozen = [(off, n) for p, off, n in zip(owner._param_list(), owner._offsets, owner._sizes) if not p.requires_grad]
        self._step += 1
        self._grad_mult_now = float(self.grad_mult) * float(getattr(owner, "_grad_mult", 1.0))
        owner._grad_mult = 1.0                      # consumed: the next backward produces freshly scaled gradients
        p16 = owner._flat16 if owner.precision == "bf16" else None
        b1, b2 = grp["betas"]
        if not frozen:
            self._launch(flat, g, p16, 0, flat.numel(), grp, b1, b2, sumsq, max_norm)
        else:
            live = _complement(frozen, flat.numel())
            for lo, hi in live:
                self._launch(flat, g, p16, lo, hi, grp, b1, b2, sumsq, max_norm)
        if p16 is not None:
            owner.mark_shadow_fresh()
        return loss

    def _step_nvls(self, nvls, owner, grp, loss):
        """Data parallel over NVSwitch multicast (dp.NvlsStep): reduce-scatter (if clip_grad_norm_ has not done it yet) -> Adam
        on this rank's slice -> all-gather of the updated masters and bf16 shadow, all inside two kernels."""
        if any(not p.requires_grad for p in owner._param_list()):
            raise RuntimeError("DataParallel(mode='nvls') does not support frozen parameters; use mode='nccl'")
        max_norm = 0.0
        if owner._pending_clip is not None:
            max_norm = owner._pending_clip[1]
            owner._pending_clip = None
        elif self.max_grad_norm is not None:
            max_norm = float(self.max_grad_norm)
        nvls.reduce()
        self._step += 1
        mult = float(self.grad_mult) * float(getattr(owner, "_grad_mult", 1.0))
        owner._grad_mult = 1.0
        nvls.adam(self, grp, self._m, self._v, self._step, mult, max_norm)
        if owner._flat16 is not None:
            owner.mark_shadow_fresh()
        return loss

    def _sync_lr(self, non_blocking: bool = True):
        """capturable mode: hand the current param_group lr (set by an LR scheduler on the host) to the device."""
        lr = float(self.param_groups[0]["lr"])
        if float(self._lr_host[0]) != lr:
            self._lr_host[0] = lr
            self._lr_dev.copy_(self._lr_host, non_blocking=non_blocking)

    def _launch(self, flat, g, p16, lo, hi, grp, b1, b2, sumsq, max_norm):
        with L.nvtx_range("vitk/adam_step"):
            self._launch_range(flat, g, p16, lo, hi, grp, b1, b2, sumsq, max_norm)

    def _launch_range(self, flat, g, p16, lo, hi, grp, b1, b2, sumsq, max_norm):
        n = hi - lo
        if n <= 0:
            return
        if self.capturable:
            if lo != 0 or hi != flat.numel():
                raise RuntimeError("FusedAdam(capturable=True) does not support frozen parameter ranges")
            if not torch.cuda.is_current_stream_capturing():
                self._sync_lr()
            L.call("vitk_adam_step_graph", flat.data_ptr(), g.data_ptr(), self._m.data_ptr(), self._v.data_ptr(),
                   p16.data_ptr() if p16 is not None else None, n, L.ptr(self._lr_dev), L.ptr(self._step_dev), L.ptr(self._hyper),
                   float(b1), float(b2), float(grp["eps"]), float(grp["weight_decay"]), 1 if grp["adamw"] else 0,
                   float(self._grad_mult_now), L.ptr(sumsq), float(max_norm), L.stream_ptr())
            return
        L.call("vitk_adam_step", flat.data_ptr() + 4 * lo, g.data_ptr() + 4 * lo, self._m.data_ptr() + 4 * lo,
               self._v.data_ptr() + 4 * lo, (p16.data_ptr() + 2 * lo) if p16 is not None else None, n,
               float(grp["lr"]), float(b1), float(b2), float(grp["eps"]), float(grp["weight_decay"]),
               1 if grp["adamw"] else 0, self._step, float(self._grad_mult_now), L.ptr(sumsq), float(max_norm), L.stream_ptr())

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none=set_to_none)
        self._owner._gathered_serial = None   # .grad tensors changed: the next clip / step re-validates them

    # torch.optim.AdamW-compatible state layout so save_checkpoint (train_advanced.py:475-489) round-trips
    def _full_moments(self):
        """(m, v) over the whole flat index space.  NVLS data parallel keeps only this rank's slice: gather the others."""
        nvls = getattr(self._owner, "_nvls", None)
        if nvls is None:
            return self._m, self._v
        import torch.distributed as dist
        dev = self._m[0].device
        fm = torch.zeros(nvls.total, dtype=torch.float32, device=dev)
        fv = torch.zeros_like(fm)
        for k, d in enumerate(nvls.domains):      # every rank's slice of the domain, padded to the common slice length
            for src, dst in ((self._m[k], fm), (self._v[k], fv)):
                mine = torch.zeros(d["per"], dtype=torch.float32, device=dev)
                mine[:d["n"]].copy_(src[:d["n"]])
                allp = torch.empty(d["per"] * nvls.world, dtype=torch.float32, device=dev)
                dist.all_gather_into_tensor(allp, mine, group=nvls.group)
                n = d["hi"] - d["lo"]
                dst[d["lo"]:d["hi"]].copy_(allp[:n])
        return fm, fv

    def state_dict(self):
        owner = self._owner
        state = {}
        if self._m is not None:
            fm, fv = self._full_moments()
            for i, (p, off, n) in enumerate(zip(owner._param_list(), owner._offsets, owner._sizes)):
                state[i] = {"step": torch.tensor(float(self._step)),
                            "exp_avg": fm[off:off + n].view(p.shape).clone(),
                            "exp_avg_sq": fv[off:off + n].view(p.shape).clone()}
        grp = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        grp["params"] = list(range(len(owner._param_list())))
        return {"state": state, "param_groups": [grp]}

    def load_state_dict(self, sd):
        owner = self._owner
        flat = owner.flat_params()
        self._ensure_state(flat)
        for k, v in sd["param_groups"][0].items():
            if k != "params":
                self.param_groups[0][k] = v
        nvls = getattr(owner, "_nvls", None)
        fm, fv = (self._m, self._v) if nvls is None else (torch.zeros(nvls.total, dtype=torch.float32, device=flat.device),
                                                          torch.zeros(nvls.total, dtype=torch.float32, device=flat.device))
        for i, (p, off, n) in enumerate(zip(owner._param_list(), owner._offsets, owner._sizes)):
            st = sd["state"].get(i)
            if st is None:
                continue
            fm[off:off + n].copy_(st["exp_avg"].reshape(-1))
            fv[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            self._step = int(float(st["step"]))
        if self._step_dev is not None:
            self._step_dev.fill_(self._step)
        if nvls is not None:      # keep this rank's slices
            for k, d in enumerate(nvls.domains):
                self._m[k][:d["n"]].copy_(fm[d["a"]:d["a"] + d["n"]])
                self._v[k][:d["n"]].copy_(fv[d["a"]:d["a"] + d["n"]])


class FusedGradScaler:
    """``torch.cuda.amp.GradScaler`` for the fused path (the reference builds one at train_advanced.py:609 and drives it at
    :330-336: ``scaler.scale(loss).backward(); scaler.unscale_(optimizer); clip_grad_norm_(...); scaler.step(optimizer);
    scaler.update()``).  Same API and state machine -- skip the step when a gradient is non-finite, halve the scale, double
    it after ``growth_interval`` clean steps.

    ``unscale_`` divides the gradients in place (one pass over the flat buffer), so whatever follows it -- the stock
    ``torch.nn.utils.clip_grad_norm_`` of the reference loop or this package's -- sees true gradients.
    ``lazy_unscale=True`` is the fully fused variant: ``unscale_`` touches no memory, 1/scale is folded into the single
    sum-of-squares + Adam pass; it is only correct with THIS package's ``clip_grad_norm_`` and ``FusedAdam`` and raises
    otherwise.  The finiteness test is the finiteness of one sum of squares; like torch's, ``step`` reads one scalar back
    (one host sync per step; the reference loop has two more at :345-346).  bf16 needs no loss scaling; this exists so the
    reference's loop runs unmodified."""

    def __init__(self, init_scale: float = 2.0 ** 16, growth_factor: float = 2.0, backoff_factor: float = 0.5,
                 growth_interval: int = 2000, enabled: bool = True, lazy_unscale: bool = False):
        self._enabled = bool(enabled)
        self._scale = float(init_scale)
        self._growth_factor, self._backoff_factor = float(growth_factor), float(backoff_factor)
        self._growth_interval = int(growth_interval)
        self._growth_tracker = 0
        self._found_inf = False
        self._unscaled = False
        self._lazy = bool(lazy_unscale)

    def is_enabled(self):
        return self._enabled

    def get_scale(self):
        return self._scale if self._enabled else 1.0

    def scale(self, outputs):
        return outputs * self._scale if self._enabled else outputs

    @staticmethod
    def _owner(optimizer):
        owner = getattr(optimizer, "_owner", None)
        if owner is None:
            owner, _ = _owner_of(optimizer.param_groups[0]["params"][:1])
        return owner

    def unscale_(self, optimizer):
        if not self._enabled:
            return
        if self._unscaled:
            raise RuntimeError("unscale_() has already been called on this optimizer since the last update().")
        owner = self._owner(optimizer)
        if self._lazy:
            if not isinstance(optimizer, FusedAdam):
                raise TypeError("FusedGradScaler(lazy_unscale=True) needs FusedAdam: the factor is applied inside its pass")
            owner._grad_mult = 1.0 / self._scale
        else:
            g = _gather_flat_grads(owner, None)
            _scale_grads_now(owner, g, 1.0 / self._scale, None, 0.0)
        self._unscaled = True

    def step(self, optimizer, *args, **kwargs):
        if not self._enabled:
            return optimizer.step(*args, **kwargs)
        owner = self._owner(optimizer)
        if not self._unscaled:
            self.unscale_(optimizer)
        if owner._pending_clip is not None:
            sumsq = owner._pending_clip[0]           # this package's clip_grad_norm_ ran after unscale_: reuse its reduction
        else:
            # none, or the stock torch clip (which rescales in place: non-finite values stay non-finite)
            sumsq = _sumsq(owner, _gather_flat_grads(owner, None))
            if self._lazy:
                owner._pending_clip = (sumsq, 0.0)
        self._found_inf = not bool(torch.isfinite(sumsq).item())
        if self._found_inf:
            owner._pending_clip = None
            owner._grad_mult = 1.0
            return None
        return optimizer.step(*args, **kwargs)

    def update(self, new_scale=None):
        if not self._enabled:
            return
        if new_scale is not None:
            self._scale = float(new_scale)
        elif self._found_inf:
            self._scale *= self._backoff_factor
            self._growth_tracker = 0
        else:
            self._growth_tracker += 1
            if self._growth_tracker == self._growth_interval:
                self._scale *= self._growth_factor
                self._growth_tracker = 0
        self._found_inf = False
        self._unscaled = False

    def state_dict(self):
        return {"scale": self._scale, "growth_factor": self._growth_factor, "backoff_factor": self._backoff_factor,
                "growth_interval": self._growth_interval, "_growth_tracker": self._growth_tracker} if self._enabled else {}

    def load_state_dict(self, sd):
        if not sd:
            return
        self._scale = float(sd["scale"])
        self._growth_factor, self._backoff_factor = float(sd["growth_factor"]), float(sd["backoff_factor"])
        self._growth_interval, self._growth_tracker = int(sd["growth_interval"]), int(sd["_growth_tracker"])


def _complement(ranges, total):
    out, cur = [], 0
    for off, n in sorted(ranges):
        if off > cur:
            out.append((cur, off))
        cur = max(cur, off + n)
    if cur < total:
        out.append((cur, total))
    # launches need 16-byte aligned starts: offsets are multiples of 64 elements by construction
    return out


class GraphedTrainStep:
    """The reference's optimisation step (train_advanced.py:322-346: forward, focal loss, backward, clip_grad_norm_, optimizer
    step, zero_grad) captured ONCE into a CUDA graph and replayed: ~290 kernel launches and the autograd / ctypes walk of a step
    become one graph launch, so the host side of a step costs tens of microseconds instead of ~4 ms.

        step = GraphedTrainStep(model, criterion, optimizer, images, labels, max_grad_norm=1.0)
        for images, labels in loader:
            loss, metrics = step(images, labels)       # static tensors: read (or copy) them before the next call
            scheduler.step()                           # the new lr reaches the device before the next replay

    ``optimizer`` must be ``FusedAdam(capturable=True)``; batch shape and dtype are fixed by the example batch.  Construction
    runs warm-up steps and the capture on the example batch and then restores parameters, moments and step count: it has no
    training side effect.  Single-process path (data-parallel modes use the eager step)."""

    def __init__(self, model, criterion, optimizer, images, labels, max_grad_norm=1.0, warmup: int = 3):
        if not isinstance(optimizer, FusedAdam) or not optimizer.capturable:
            raise TypeError("GraphedTrainStep needs FusedAdam(capturable=True)")
        owner = optimizer._owner
        if getattr(owner, "_nvls", None) is not None or owner._bucket_hook is not None:
            raise RuntimeError("GraphedTrainStep is the single-process path; data-parallel modes use the eager step")
        self.model, self.criterion, self.opt, self.max_grad_norm = model, criterion, optimizer, max_grad_norm
        self.static_x = images.detach().clone()
        self.static_y = labels.detach().clone()
        owner._ensure_flat()
        optimizer._ensure_state(owner._flat)
        optimizer._sync_lr(non_blocking=False)
        snap = (owner._flat.clone(), optimizer._m.clone(), optimizer._v.clone(), optimizer._step)
        side = torch.cuda.Stream(device=images.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.metrics, self.grad_norm = self._eager()
        # undo the training effect of warm-up and capture
        owner._flat.copy_(snap[0]); optimizer._m.copy_(snap[1]); optimizer._v.copy_(snap[2])
        optimizer._step = snap[3]
        optimizer._step_dev.fill_(snap[3])
        owner.invalidate_shadow()
        owner._ensure_shadow(owner._param_list())
        torch.cuda.synchronize()

    def _eager(self):
        out = self.model(self.static_x)
        loss, met = self.criterion(out, self.static_y, with_metrics=True)
        loss.backward()
        norm = clip_grad_norm_(self.model.parameters(), self.max_grad_norm) if self.max_grad_norm else None
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        return loss, met, norm

    def __call__(self, images, labels):
        self.static_x.copy_(images, non_blocking=True)
        self.static_y.copy_(labels, non_blocking=True)
        self.opt._sync_lr()
        self.graph.replay()
        self.opt._step += 1
        return self.loss, self.metrics
