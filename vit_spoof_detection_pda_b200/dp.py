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


class DataParallel(nn.Module):
    """``DataParallel(ViTFaceAntiSpoofing(...).cuda())``: identical replicas, local batch per rank,
    gradients averaged over ranks during backward."""

    def __init__(self, module: nn.Module, process_group=None, bucket_mb: float = 50.0, broadcast: bool = True,
                 comm_sms: Optional[int] = None, grad_comm_dtype: Optional[torch.dtype] = None):
        super().__init__()
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (one process per GPU, backend 'nccl')")
        self.module = module
        self.group = process_group
        self.bucketer = GradBucketer(int(bucket_mb * 1e6 / 4), process_group, comm_dtype=grad_comm_dtype)
        # SMs left to the NCCL all-reduce kernels while backward runs.  The persistent GEMM CTAs need a whole SM
        # each (~225 KB smem); if NCCL's CTAs hold some SMs the GEMM grid must shrink by that many, or the CTAs
        # that cannot become resident stall their tiles until the collective ends.  Set NCCL_MAX_CTAS (before
        # init_process_group) to the same number so NCCL does not take more.
        if comm_sms is None:
            comm_sms = int(os.environ.get("NCCL_MAX_CTAS", "16"))
        self.comm_sms = comm_sms if dist.get_world_size(process_group) > 1 else 0
        if broadcast:
            flat = module.flat_params()
            dist.broadcast(flat, src=0, group=process_group)
            module.invalidate_shadow()   # the bf16 shadow (if a forward already built one) no longer matches the masters
        module._bucket_hook = self._on_stage
        module._finish_hook = self._on_finish
        self._first = True
        self._n_sms = None

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
