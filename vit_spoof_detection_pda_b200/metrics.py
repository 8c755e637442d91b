"""Evaluation post-processing on the device.

Mirrors ``find_optimal_threshold(labels, probs, config)`` of the reference (/root/reference/train_advanced.py:239-275)
and the confusion counts of ``calculate_metrics`` (/root/reference/test.py:241-243), but the scores never leave the GPU:
the validation loop calls :meth:`ThresholdSweep.update` once per batch with the device tensors ``probs[:, 1]`` and
``labels`` (train_advanced.py:387-394 appends them to host lists instead), and one ``[steps, 4]`` count tensor is read
back per epoch.  Counts are exact integers, so accuracy / precision / recall / F1 equal sklearn's bit for bit.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L


class ThresholdSweep:
    """Streaming (tp, fp, tn, fn) per decision threshold.  ``config`` needs ``threshold_min/max/steps`` like the
    reference's ``Config`` (train_advanced.py:76-79); live = 1 is the positive class."""

    def __init__(self, config=None, threshold_min: float = 0.3, threshold_max: float = 0.7, threshold_steps: int = 41,
                 device="cuda"):
        if config is not None:
            threshold_min, threshold_max, threshold_steps = config.threshold_min, config.threshold_max, config.threshold_steps
        self.thresholds = np.linspace(threshold_min, threshold_max, threshold_steps)     # float64, as the reference
        if np.any(np.diff(self.thresholds) <= 0):
            raise ValueError("thresholds must ascend")
        self.steps = int(threshold_steps)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ThresholdSweep runs on CUDA only (no CPU fallback)")
        self._th = torch.from_numpy(self.thresholds).to(self.device)
        self._hist = torch.zeros(2 * (self.steps + 1), dtype=torch.int64, device=self.device)

    def reset(self):
        self._hist.zero_()

    @torch.no_grad()
    def update(self, probs, labels):
        """probs: P(live) per sample (any float dtype, device tensor); labels: int tensor, live = 1."""
        if not (torch.is_tensor(probs) and probs.is_cuda and torch.is_tensor(labels) and labels.is_cuda):
            raise RuntimeError("ThresholdSweep.update needs CUDA tensors (no CPU fallback)")
        p = probs.detach().reshape(-1).to(torch.float32).contiguous()     # fp16 -> fp32 is exact
        y = labels.detach().reshape(-1).to(torch.int64).contiguous()
        if p.numel() != y.numel():
            raise ValueError("probs and labels differ in length")
        if p.numel() == 0:
            return
        L.call("vitk_threshold_hist", L.ptr(p), L.ptr(y), L.ptr(self._th), p.numel(), self.steps, L.ptr(self._hist),
               L.stream_ptr())

    @torch.no_grad()
    def counts(self):
        """int64 device tensor [steps, 4] = (tp, fp, tn, fn) per threshold."""
        out = torch.empty(self.steps, 4, dtype=torch.int64, device=self.device)
        L.call("vitk_threshold_counts", L.ptr(self._hist), self.steps, L.ptr(out), L.stream_ptr())
        return out

    @staticmethod
    def metrics_from_counts(tp: int, fp: int, tn: int, fn: int):
        """accuracy, precision, recall, f1 as sklearn (average='binary', zero_division=0) computes them."""
        n = tp + fp + tn + fn
        acc = (tp + tn) / n if n else 0.0
        prec = tp / (tp + fp) if (tp + fp) else 0.0
        rec = tp / (tp + fn) if (tp + fn) else 0.0
        f1 = (2 * tp / (2 * tp + fp + fn)) if (2 * tp + fp + fn) else 0.0
        return acc, prec, rec, f1

    def results(self):
        """List of {'threshold', 'accuracy', 'precision', 'recall', 'f1'} (the rows the reference logs)."""
        rows = []
        for th, (tp, fp, tn, fn) in zip(self.thresholds, self.counts().cpu().tolist()):
            acc, prec, rec, f1 = self.metrics_from_counts(tp, fp, tn, fn)
            rows.append({"threshold": float(th), "accuracy": acc, "precision": prec, "recall": rec, "f1": f1})
        return rows

    def best(self):
        """(best_threshold, best_f1, best_acc): first threshold with the strictly largest F1, defaults 0.5 / 0 / 0."""
        best_t, best_f1, best_acc = 0.5, 0, 0
        for r in self.results():
            if r["f1"] > best_f1:
                best_f1, best_t, best_acc = r["f1"], r["threshold"], r["accuracy"]
        return best_t, best_f1, best_acc


def find_optimal_threshold(labels, probs, config):
    """Drop-in for the reference function (train_advanced.py:239): same arguments, same return triple.  ``labels`` and
    ``probs`` may be device tensors (no host round trip) or host arrays (copied once)."""
    dev = probs.device if torch.is_tensor(probs) and probs.is_cuda else torch.device("cuda")
    sweep = ThresholdSweep(config, device=dev)
    p = probs if torch.is_tensor(probs) else torch.from_numpy(np.asarray(probs))
    y = labels if torch.is_tensor(labels) else torch.from_numpy(np.asarray(labels))
    sweep.update(p.to(dev), y.to(dev))
    return sweep.best()


@torch.no_grad()
def confusion_counts(labels, preds):
    """(tn, fp, fn, tp) like ``confusion_matrix(y_true, y_pred).ravel()`` (test.py:242-243) from device tensors."""
    sweep = ThresholdSweep(threshold_min=0.5, threshold_max=0.5, threshold_steps=1, device=preds.device)
    sweep.update(preds.to(torch.float32), labels)
    tp, fp, tn, fn = sweep.counts()[0].tolist()
    return tn, fp, fn, tp
