"""Batch-sharded data parallelism: one process per GPU, bucketed gradient all-reduce over NCCL
(NVLink 5 / NVSwitch) overlapped with backward.

The reference is single-GPU (no torch.distributed call site); this is the new
functionality BASELINE config #5 asks for.  Backward runs in stages (module.stage_ranges()); the flat
gradient range a stage finalises is contiguous and adjacent to the previous stage's range, so buckets
are plain slices of the flat gradient buffer -- no packing copies.  Each full bucket is all-reduced
asynchronously on the process group's communication stream while the next stage computes; the last
bucket's wait is the only exposed communication.  Inference shards by batch with no collective.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn


class GradBucketer:
    """Greedy merge of consecutive (descending, adjacent) flat ranges into buckets of >= bucket_elems,
    all-reduced (average) as soon as they close.  Device- and backend-agnostic (tested with gloo on CPU)."""

    def __init__(self, bucket_elems: int, group=None, comm_dtype: Optional[torch.dtype] = None):
        self.bucket_elems = int(bucket_elems)
        self.group = group
        # None: the fp32 slices are all-reduced in place.  torch.bfloat16 / torch.float16 (opt-in): every bucket is cast to
        # a 16-bit copy, that copy is all-reduced and written back into the fp32 slice at finish() -- the semantics of
        # torch DDP's bf16_compress_hook / fp16_compress_hook: half the bytes on NVLink for two extra HBM passes.
        self.comm_dtype = comm_dtype
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._open: Optional[Tuple[int, int]] = None
        self._works: List = []
        self.launched: List[Tuple[int, int]] = []   # for tests / introspection

    def reset(self):
        self._open = None
        self._works = []
        self.launched = []

    def _launch(self, flat: torch.Tensor, lo: int, hi: int):
        self.launched.append((lo, hi))
        if self.world == 1:
            return
        view = flat[lo:hi]
        buf = view if self.comm_dtype is None else view.to(self.comm_dtype)
        backend = dist.get_backend(self.group)
        if backend == "nccl":
            work = dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            self._works.append((work, view, buf, False))
        else:  # gloo has no AVG
            work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._works.append((work, view, buf, True))

    def add(self, flat: torch.Tensor, lo: int, hi: int):
        """Register that flat[lo:hi] is final. Ranges must arrive adjacent and descending (or be disjoint)."""
        if self._open is None:
            self._open = (lo, hi)
        elif hi == self._open[0]:
            self._open = (lo, self._open[1])
        elif lo == self._open[1]:
            self._open = (self._open[0], hi)
        else:  # not adjacent: close what we have
            self._launch(flat, *self._open)
            self._open = (lo, hi)
        if self._open[1] - self._open[0] >= self.bucket_elems:
            self._launch(flat, *self._open)
            self._open = None

    def finish(self, flat: torch.Tensor):
        if self._open is not None:
            self._launch(flat, *self._open)
            self._open = None
        for work, view, buf, divide in self._works:
            work.wait()   # stream-ordered for NCCL (no host block), blocking for gloo
            if buf is not view:
                view.copy_(buf)          # decompress: the reduced 16-bit bucket back into the fp32 gradient slice
            if divide:
                view.div_(self.world)
        self._works = []


class NvlsStep:
    """Gradient all-reduce + global-norm clip + Adam + parameter all-gather as two kernels per rank over NVLink / NVSwitch
    multicast (csrc/dp_nvls.cu).  The module's flat parameter / bf16-shadow / gradient buffers are re-allocated as symmetric
    memory (torch.distributed._symmetric_memory: the plumbing -- allocation, handle exchange, stream-ordered barriers).

    The flat index space is cut into ``n_domains`` domains along the order in which backward finalises gradients (domain 0 =
    the classifier and the last blocks, final first); rank r owns slice r of every domain.  The reduce-scatter of a domain is
    enqueued on a side stream as soon as the backward stage that completes it returns (a cross-rank barrier on that stream,
    then the kernel with a bounded grid), so only the LAST domain's reduce-scatter and the Adam + all-gather kernels run after
    backward.  ``overlap=False`` (or gradient accumulation: see below) does all of it at the step.

    With ``overlap=True`` every backward must be followed by an optimizer step: an early reduce replaces this rank's slice by
    the cross-rank average, so accumulating a second local backward on top would mix averaged and local gradients --
    ``backward`` raises in that case.  Raises at construction when the group has no multicast (NVLS) support:
    DataParallel(mode="auto") then falls back to the bucketed NCCL all-reduce."""

    def __init__(self, module, group=None, n_domains: int = 4, overlap: bool = True):
        import torch.distributed._symmetric_memory as symm
        from . import _lib as L
        self.L = L
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        module._ensure_flat()
        total = module._total
        dev = module._flat.device
        self.module = module
        self.total = total
        self.overlap = bool(overlap)

        # ---- domains: consecutive backward stages grouped to roughly equal element counts
        stages = module.stage_ranges()                   # stage s -> (lo, hi), descending and adjacent
        n_domains = max(1, min(int(n_domains), len(stages)))
        target, doms, cur_hi, acc = total / n_domains, [], stages[0][1], 0
        for s_idx, (lo, hi) in enumerate(stages):
            acc += hi - lo
            last = s_idx == len(stages) - 1
            if last or (acc >= target and len(doms) < n_domains - 1):
                doms.append({"lo": lo, "hi": cur_hi, "stage": s_idx})
                cur_hi, acc = lo, 0
        assert doms[-1]["lo"] == 0 and doms[0]["hi"] == total
        for d in doms:
            n = d["hi"] - d["lo"]
            per = (-(-n // self.world) + 3) // 4 * 4     # slice length: a multiple of 4 elements (16 bytes)
            a = min(d["hi"], d["lo"] + self.rank * per)
            b = min(d["hi"], a + per)
            d["a"], d["n"] = a, b - a                    # this rank's slice [a, a + n) of the domain
            d["per"] = per
        self.domains = doms
        self.n_slots = len(doms) * self.world
        assert self.n_slots <= 64

        def alloc(n, dtype):
            t = symm.empty(n, dtype=dtype, device=dev)
            h = symm.rendezvous(t, self.group)
            if not h.multicast_ptr:
                raise RuntimeError("symmetric memory without multicast support (no NVLS on this system)")
            t.zero_()
            return t, h

        self.p, self.hp = alloc(total, torch.float32)
        self.g, self.hg = alloc(total, torch.float32)
        self.p16, self.hp16 = alloc(total, torch.bfloat16) if module.precision == "bf16" else (None, None)
        self.table, self.ht = alloc(64, torch.float32)
        module.adopt_flat_buffers(self.p, self.p16, self.g)
        self.scratch = [torch.empty(L.load().vitk_nvls_scratch_floats(), dtype=torch.float32, device=dev) for _ in doms]
        self._reduced_serial = None
        self._early = set()          # domains already reduced (under backward) for the current backward
        self.side = torch.cuda.Stream(device=dev)        # early reduce-scatters
        self.side2 = torch.cuda.Stream(device=dev)       # background fp32-master copies
        self.early_ctas = int(os.environ.get("VITK_NVLS_EARLY_CTAS", "48"))
        # fp32 masters of the tensor-core GEMM weights (98.6 % of the parameters) are only ever read through the bf16 shadow
        # by forward / backward: the fused step multicasts their SHADOW, keeps the master update local to the owning rank, and
        # a bounded-CTA multicast copy on a side stream brings the other ranks' fp32 copies up to date under the next forward
        # pass (345 MB less on the critical path per step).  state_dict() / flat_params() wait for that copy.
        self.lazy = module.precision == "bf16" and os.environ.get("VITK_NVLS_EAGER_MASTERS") != "1"
        self.masters_done = None
        if self.lazy:
            import ctypes as C
            for d in doms:
                lo, hi = d["a"], d["a"] + d["n"]
                rs = []
                for (name, off, n) in module.param_ranges():
                    is_gemm_w = name.startswith("vit.") and name.endswith(".weight") and (
                        ".attn.qkv." in name or ".attn.proj." in name or ".mlp.fc1." in name or ".mlp.fc2." in name or "patch_embed" in name)
                    x, y = max(off, lo), min(off + n, hi)
                    if is_gemm_w and x < y:
                        rs += [x - lo, y - lo]
                assert len(rs) // 2 <= 64
                d["ranges"] = ((C.c_int64 * max(1, len(rs)))(*rs), len(rs) // 2)
        module._masters_sync = self.sync_masters
        module._bucket_hook = self._on_stage
        self._prof = [] if os.environ.get("VITK_NVLS_PROF") == "1" else None     # per-phase CUDA events (diagnostics)
        torch.cuda.synchronize(dev)
        self.hg.barrier(channel=0)

    @staticmethod
    def _mc(t, h, elem_off):
        return h.multicast_ptr + (t.data_ptr() - h.buffer_ptrs[h.rank]) + elem_off * t.element_size()

    def alloc_moments(self):
        dev = self.p.device
        return ([torch.zeros(max(4, d["n"]), dtype=torch.float32, device=dev) for d in self.domains],
                [torch.zeros(max(4, d["n"]), dtype=torch.float32, device=dev) for d in self.domains])

    def _mark(self):
        if self._prof is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self._prof.append(e)

    def report(self):
        """VITK_NVLS_PROF=1: mean milliseconds between the marks of a step (barrier, last reduce-scatter, barrier, Adam +
        all-gather kernels, barrier) over the second half of the recorded steps."""
        if not self._prof:
            return None
        torch.cuda.synchronize()
        k = 6
        n = len(self._prof) // k
        names = ["barrier_grads_final", "reduce_scatter_after_backward", "barrier_table", "adam_allgather", "barrier_params"]
        tot = [0.0] * (k - 1)
        for i in range(n // 2, n):
            ev = self._prof[k * i:k * i + k]
            for j in range(k - 1):
                tot[j] += ev[j].elapsed_time(ev[j + 1])
        m = max(1, n - n // 2)
        return {names[j]: round(tot[j] / m, 4) for j in range(k - 1)}

    def sync_masters(self):
        """Host-wait until every rank's background copy of its fp32 master slices has landed everywhere (the event follows a
        cross-rank barrier on the side stream): after this, this rank's fp32 parameters are complete and current."""
        if self.lazy and self.masters_done is not None:
            self.masters_done.synchronize()

    def _reduce_domain(self, k, stream, max_ctas):
        d, L = self.domains[k], self.L
        if d["n"] > 0:
            L.call("vitk_nvls_reduce_sumsq", self.g.data_ptr() + 4 * d["a"], self._mc(self.g, self.hg, d["a"]), d["n"],
                   1.0 / self.world, L.ptr(self.scratch[k]), self._mc(self.table, self.ht, 0), k * self.world + self.rank,
                   max_ctas, stream.cuda_stream)

    def _on_stage(self, stage, lo, hi, flat_grad):
        """backward hook (module._bucket_hook): the stage that completes a domain has been enqueued -> start that domain's
        reduce-scatter on the side stream (barrier: the domain is final on every rank; then the kernel, bounded grid)."""
        accumulating = flat_grad.data_ptr() != self.g.data_ptr()     # gradients go to the second buffer, autograd adds them to .grad
        if stage == 0:
            if accumulating and self._early:
                raise RuntimeError("DataParallel(mode='nvls', overlap=True): a second backward before the optimizer step "
                                   "(gradient accumulation) would mix averaged and local gradients; use nvls_overlap=False")
            if not accumulating:
                self._early = set()      # a fresh backward rewrites the flat buffer: nothing of it is reduced yet
        if not self.overlap or accumulating:
            return                       # everything is reduced at the step
        for k, d in enumerate(self.domains[:-1]):
            if d["stage"] == stage:
                main = torch.cuda.current_stream()
                self.side.wait_stream(main)
                with torch.cuda.stream(self.side):
                    self.hg.barrier(channel=1 + k)
                    self._reduce_domain(k, self.side, self.early_ctas)
                self._early.add(k)

    def reduce(self):
        """The reduce-scatters not yet done under backward (at least the last domain's) + barriers: afterwards this rank's
        slices hold the averaged gradients and every rank's norm table is complete.  Once per backward."""
        serial = getattr(self.module, "_bwd_serial", None)
        if serial is not None and self._reduced_serial == serial:
            return
        self._reduced_serial = serial
        main = torch.cuda.current_stream()
        if self.lazy and self.masters_done is not None:
            main.wait_event(self.masters_done)   # the previous step's background master copy is complete
        self._mark()
        self.hg.barrier(channel=0)               # every rank's gradients are final
        self._mark()
        for k in range(len(self.domains)):
            if k not in self._early:
                self._reduce_domain(k, main, 0)
        main.wait_stream(self.side)              # the early reduce-scatters of this rank
        self._early = set()
        self._mark()
        self.hg.barrier(channel=0)               # ... and of every other rank: the norm table is complete
        self._mark()

    def total_norm(self, grad_mult: float = 1.0):
        n = self.table[:self.n_slots].sum().sqrt().reshape(())
        return n * grad_mult if grad_mult != 1.0 else n

    def adam(self, opt, grp, ms, vs, step, grad_mult, max_norm):
        L = self.L
        b1, b2 = grp["betas"]
        st = L.stream_ptr()
        for k, d in enumerate(self.domains):
            if d["n"] <= 0:
                continue
            rg = d.get("ranges") if self.lazy else None
            L.call("vitk_nvls_adam_bcast", self.p.data_ptr() + 4 * d["a"], self._mc(self.p, self.hp, d["a"]),
                   self._mc(self.p16, self.hp16, d["a"]) if self.p16 is not None else None, self.g.data_ptr() + 4 * d["a"],
                   L.ptr(ms[k]), L.ptr(vs[k]), d["n"], float(grp["lr"]), float(b1), float(b2), float(grp["eps"]),
                   float(grp["weight_decay"]), 1 if grp["adamw"] else 0, step, float(grad_mult), L.ptr(self.table), self.n_slots,
                   float(max_norm), rg[0] if rg else None, rg[1] if rg else 0, st)
        self._mark()
        self.hg.barrier(channel=0)      # every rank's slices of the bf16 shadow (and of the eagerly sent masters) have landed everywhere
        self._mark()
        if self.lazy:
            main = torch.cuda.current_stream()
            self.side2.wait_stream(main)
            with torch.cuda.stream(self.side2):
                for d in self.domains:
                    if d["n"] > 0:
                        L.call("vitk_nvls_bcast_f32", self.p.data_ptr() + 4 * d["a"], self._mc(self.p, self.hp, d["a"]), d["n"],
                               int(os.environ.get("VITK_NVLS_BCAST_CTAS", "32")), self.side2.cuda_stream)
                self.hp.barrier(channel=1)
                ev = torch.cuda.Event()
                ev.record(self.side2)
            self.masters_done = ev


class DataParallel(nn.Module):
    """``DataParallel(ViTFaceAntiSpoofing(...).cuda())``: identical replicas, local batch per rank,
    gradients averaged over ranks during backward."""

    def __init__(self, module: nn.Module, process_group=None, bucket_mb: float = 50.0, broadcast: bool = True,
                 comm_sms: Optional[int] = None, grad_comm_dtype: Optional[torch.dtype] = None, mode: str = "auto",
                 nvls_overlap: Optional[bool] = None, nvls_domains: Optional[int] = None):
        """mode: "nvls" = the optimizer step fused with its collectives over NVSwitch multicast (NvlsStep; needs FusedAdam and
        NVLS); "nccl" = bucketed NCCL all-reduce overlapped with backward; "auto" = nvls when available (CUDA, NCCL backend,
        multicast support, no frozen parameters), else nccl.  nvls_overlap (default True): reduce-scatter the gradient domains
        under the backward pass -- every backward must then be followed by an optimizer step; False: everything at the step
        (gradient accumulation allowed).  nvls_domains: number of gradient domains (default 4)."""
        super().__init__()
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (one process per GPU, backend 'nccl')")
        if mode not in ("auto", "nvls", "nccl"):
            raise ValueError(mode)
        self.module = module
        self.group = process_group
        world = dist.get_world_size(process_group)
        self.bucketer = GradBucketer(int(bucket_mb * 1e6 / 4), process_group, comm_dtype=grad_comm_dtype)
        # SMs left to the NCCL all-reduce kernels while backward runs.  The persistent GEMM CTAs need a whole SM
        # each (~225 KB smem); if NCCL's CTAs hold some SMs the GEMM grid must shrink by that many, or the CTAs
        # that cannot become resident stall their tiles until the collective ends.  Set NCCL_MAX_CTAS (before
        # init_process_group) to the same number so NCCL does not take more.
        if comm_sms is None:
            comm_sms = int(os.environ.get("NCCL_MAX_CTAS", "16"))
        self.comm_sms = comm_sms if world > 1 else 0
        if broadcast:
            flat = module.flat_params()
            dist.broadcast(flat, src=0, group=process_group)
            module.invalidate_shadow()   # the bf16 shadow (if a forward already built one) no longer matches the masters
        self._first = True
        self._n_sms = None
        self.mode = "nccl"
        want_nvls = mode == "nvls" or (mode == "auto" and world > 1 and dist.get_backend(process_group) == "nccl"
                                       and all(p.requires_grad for p in module.parameters()))
        if want_nvls and world > 1:
            try:
                if nvls_domains is None:
                    nvls_domains = int(os.environ.get("VITK_NVLS_DOMAINS", "6"))
                if nvls_overlap is None:
                    nvls_overlap = os.environ.get("VITK_NVLS_OVERLAP", "1") != "0"
                module._nvls = NvlsStep(module, process_group, n_domains=nvls_domains, overlap=nvls_overlap)
                self.mode = "nvls"
            except Exception as e:  # noqa: BLE001 -- no multicast / no symmetric memory on this system
                if mode == "nvls":
                    raise
                import warnings
                warnings.warn(f"DataParallel: NVLS step unavailable ({e!r}); using the bucketed NCCL all-reduce")
        if self.mode == "nccl":
            module._bucket_hook = self._on_stage
            module._finish_hook = self._on_finish

    def _on_stage(self, stage: int, lo: int, hi: int, flat_grad: torch.Tensor):
        if stage == 0:
            self.bucketer.reset()
        self.bucketer.add(flat_grad, lo, hi)
        if self.comm_sms and self.bucketer._works:
            # a collective is (or may still be) in flight: later stages size their persistent grids for the SMs left
            self.module._sm_budget = max(1, self._hw_sms() - self.comm_sms)

    def _hw_sms(self) -> int:
        if self._n_sms is None:
            import ctypes as C
            from . import _lib as L
            n = C.c_int(0)
            L.load().vitk_device_info(C.byref(n), None, None)
            self._n_sms = n.value
        return self._n_sms

    def _on_finish(self, flat_grad: torch.Tensor):
        self.bucketer.finish(flat_grad)
        self.module._sm_budget = 0

    def forward(self, x):
        return self.module(x)

    def state_dict(self, *a, **k):
        return self.module.state_dict(*a, **k)

    def load_state_dict(self, *a, **k):
        return self.module.load_state_dict(*a, **k)


class DevicePrefetcher:
    """Host->device input pipeline for the train / eval loops (the reference's
    ``images.to(device, non_blocking=True)`` at train_advanced.py:323-324, made asynchronous): batch i+1 is
    copied from pinned host memory on a side stream while step i computes, so the H2D transfer (38.5 MB per
    64-image fp32 batch, ~1.5 ms over PCIe 5) leaves the critical path.  Iterate it like the loader::

        for images, labels in DevicePrefetcher(loader, device): ...

    Buffer-reuse contract: the yielded tensors are two preallocated device slots; a slot is overwritten two iterations
    later.  Consumers that keep a batch's tensors beyond the next iteration (the reference's eval loops collect
    ``labels`` in lists before ``.cpu()``) must copy them, or construct the prefetcher with ``clone=True``.
    """

    def __init__(self, loader, device, clone: bool = False):
        self.loader = loader
        self.clone = bool(clone)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._bufs = [None, None]          # two preallocated device slots (no allocator traffic in steady state)
        self._free = [None, None]          # compute-stream event after which slot k may be overwritten

    def _issue(self, batch, slot):
        dst = self._bufs[slot]
        if dst is None or any(torch.is_tensor(s) and (d is None or d.shape != s.shape or d.dtype != s.dtype)
                              for s, d in zip(batch, dst)) or len(dst) != len(batch):
            dst = [torch.empty(s.shape, dtype=s.dtype, device=self.device) if torch.is_tensor(s) else None for s in batch]
            self._bufs[slot] = dst
            # fresh memory from the caching allocator may be a block whose last user is still queued on the compute
            # stream: the copy stream must not write it before that work has drained
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
        if self._free[slot] is not None:
            self.stream.wait_event(self._free[slot])   # the step that last read this slot has finished on the GPU
        with torch.cuda.stream(self.stream):
            for s, d in zip(batch, dst):
                if torch.is_tensor(s):
                    d.copy_(s, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(self.stream)
        out = tuple(d if torch.is_tensor(s) else s for s, d in zip(batch, dst))
        return out, ev

    def __iter__(self):
        it = iter(self.loader)
        slot = 0
        try:
            nxt = self._issue(next(it), slot)
        except StopIteration:
            return
        while nxt is not None:
            cur, ev = nxt
            cur_slot = slot
            slot ^= 1
            try:
                nxt = self._issue(next(it), slot)      # copy of batch i+1 overlaps the compute of batch i
            except StopIteration:
                nxt = None
            compute = torch.cuda.current_stream(self.device)
            compute.wait_event(ev)
            yield tuple(t.clone() if torch.is_tensor(t) else t for t in cur) if self.clone else cur
            # everything the consumer enqueued for this batch is now in the compute stream: mark the slot reusable
            done = torch.cuda.Event()
            done.record(compute)
            self._free[cur_slot] = done


class HostScalars:
    """Device->host read-back of per-step scalars (loss, accuracy) that does not stall the compute stream.

    The reference reads ``loss.item()`` / ``acc.item()`` right after enqueueing the optimizer step
    (train_advanced.py:345-346): the copy is stream-ordered behind backward and Adam, the host blocks until the whole
    step has drained and the GPU then idles while the next step is enqueued (measured here: +0.76 ms on a 9.3 ms step).
    ``push(*tensors)`` instead copies the 0-dim tensors into pinned host memory on a side stream as soon as THEY are
    ready (an event recorded right after the loss kernel) and returns the values pushed one call earlier -- exact,
    one step late, no compute-stream sync; ``flush()`` returns the last ones."""

    def __init__(self, device, slots: int = 4, width: int = 4):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._host = torch.empty(slots, width, dtype=torch.float64).pin_memory()
        self._stage = torch.empty(slots, width, dtype=torch.float64, device=self.device)
        self._pending = []           # (slot, n, event)
        self._next = 0

    def _read(self):
        slot, n, ev = self._pending.pop(0)
        ev.synchronize()
        return tuple(self._host[slot, :n].tolist())

    @torch.no_grad()
    def push(self, *tensors):
        out = self._read() if self._pending else None
        slot = self._next
        self._next = (self._next + 1) % self._host.shape[0]
        n = len(tensors)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))       # everything the scalars depend on is enqueued
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            for i, t in enumerate(tensors):
                self._stage[slot, i].copy_(t.detach().reshape(()).to(torch.float64), non_blocking=True)
            self._host[slot, :n].copy_(self._stage[slot, :n], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        self._pending.append((slot, n, done))
        return out

    def flush(self):
        out = None
        while self._pending:
            out = self._read()
        return out
