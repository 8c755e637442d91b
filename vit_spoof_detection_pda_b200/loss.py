"""FocalLoss drop-in (reference: /root/reference/train_advanced.py:90-107) on the fused vitk kernel.

``alpha`` may be the reference's scalar or a per-class sequence (class-weighted focal: the class weights
the reference computes at train_advanced.py:521-529 but only feeds to weighted CE).  One kernel launch
produces the loss, d loss / d logits, softmax P(live) = probs[:, 1], argmax predictions and the number of
correct predictions (train_advanced.py:342-343, 387-394; test.py:212-217).
"""
from __future__ import annotations

from typing import Sequence, Union

import torch
import torch.nn as nn

from . import _lib as L

_RED = {"mean": 0, "sum": 1, "none": 2}


class _FocalFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, alpha_t, gamma, reduction, grad_scale, want_metrics):
        if not logits.is_cuda:
            raise RuntimeError("vitk FocalLoss needs CUDA tensors (no CPU fallback)")
        z = logits.detach().to(torch.float32).contiguous()
        t = targets.detach().to(torch.int64).contiguous()
        B, Cn = z.shape
        per = torch.empty(B, dtype=torch.float32, device=z.device)
        out = torch.empty(1, dtype=torch.float32, device=z.device)
        dz = torch.empty_like(z)
        probs1 = preds = ncorrect = None
        if want_metrics:
            probs1 = torch.empty(B, dtype=torch.float32, device=z.device)
            preds = torch.empty(B, dtype=torch.int64, device=z.device)
            ncorrect = torch.empty(1, dtype=torch.int32, device=z.device)
        L.call("vitk_focal_fwd_bwd", L.ptr(z), L.ptr(t), L.ptr(alpha_t), float(gamma), reduction, float(grad_scale),
               L.ptr(per), L.ptr(out), L.ptr(dz), L.ptr(probs1), L.ptr(preds), L.ptr(ncorrect), B, Cn, L.stream_ptr())
        ctx.save_for_backward(dz)
        ctx.reduction = reduction
        ctx.in_dtype = logits.dtype
        ctx.metrics = (probs1, preds, ncorrect)
        res = per if reduction == 2 else out.reshape(())
        ctx.mark_non_differentiable(*[m for m in (probs1, preds, ncorrect) if m is not None])
        if want_metrics:
            return res, probs1, preds, ncorrect
        return res

    @staticmethod
    def backward(ctx, gout, *_):
        (dz,) = ctx.saved_tensors
        g = dz * (gout.reshape(-1, 1) if ctx.reduction == 2 else gout)
        return g.to(ctx.in_dtype), None, None, None, None, None, None


class FocalLoss(nn.Module):
    """``FocalLoss(alpha=0.25, gamma=2.0, reduction='mean')(inputs[B,C], targets[B] int64)``."""

    def __init__(self, alpha: Union[float, Sequence[float]] = 0.25, gamma: float = 2.0, reduction: str = "mean",
                 grad_scale: float = 1.0):
        super().__init__()
        if reduction not in _RED:
            raise ValueError(reduction)
        self.alpha = alpha
        self.gamma = gamma
        self.reduction = reduction
        self.grad_scale = grad_scale   # extra factor on d loss / d logits (1/world for data parallel)
        self._alpha_cache = {}
        self.last_metrics = None

    def _alpha_tensor(self, num_classes, device):
        key = (num_classes, str(device), str(self.alpha))
        t = self._alpha_cache.get(key)
        if t is None:
            if isinstance(self.alpha, (int, float)):
                vals = [float(self.alpha)] * num_classes
            else:
                vals = [float(a) for a in self.alpha]
                if len(vals) != num_classes:
                    raise ValueError("per-class alpha must have num_classes entries")
            t = torch.tensor(vals, dtype=torch.float32, device=device)
            self._alpha_cache = {key: t}
        return t

    def forward(self, inputs, targets, with_metrics: bool = False):
        a = self._alpha_tensor(inputs.shape[1], inputs.device)
        out = _FocalFunction.apply(inputs, targets, a, self.gamma, _RED[self.reduction], self.grad_scale, with_metrics)
        if with_metrics:
            loss, probs1, preds, ncorrect = out
            self.last_metrics = {"probs_live": probs1, "preds": preds, "ncorrect": ncorrect}
            return loss, self.last_metrics
        return out


@torch.no_grad()
def eval_postprocess(logits):
    """softmax P(live) = probs[:,1] and argmax on device in one launch (test.py:212-217)."""
    B, Cn = logits.shape
    z = logits.detach().to(torch.float32).contiguous()
    dummy_t = torch.zeros(B, dtype=torch.int64, device=z.device)
    alpha = torch.ones(Cn, dtype=torch.float32, device=z.device)
    per = torch.empty(B, dtype=torch.float32, device=z.device)
    probs1 = torch.empty(B, dtype=torch.float32, device=z.device)
    preds = torch.empty(B, dtype=torch.int64, device=z.device)
    L.call("vitk_focal_fwd_bwd", L.ptr(z), L.ptr(dummy_t), L.ptr(alpha), 2.0, 2, 1.0, L.ptr(per), None, None,
           L.ptr(probs1), L.ptr(preds), None, B, Cn, L.stream_ptr())
    return probs1, preds
