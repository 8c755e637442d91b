"""CPU restatement (numpy) of the reference's evaluation post-processing -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(vit_spoof_detection_pda_b200/) never does.

  threshold_sweep_counts   the 41-point decision-threshold sweep of find_optimal_threshold
                           (/root/reference/train_advanced.py:239-275): preds = (probs >= thresh), then
                           accuracy / precision / recall / F1 (sklearn, average='binary', zero_division=0) per
                           threshold and the first threshold with the strictly largest F1
  confusion_counts         tn, fp, fn, tp of calculate_metrics (/root/reference/test.py:241-243,
                           train_advanced.py: validate) for given predictions
Labels: live = 1 (positive), spoof = 0 (train_advanced.py:155,160).  Pinned against the reference functions themselves:
oracle/make_golden_eval.py runs them here and stores tests/golden/threshold_golden.json.
"""
from __future__ import annotations

import numpy as np


def thresholds_of(t_min: float, t_max: float, steps: int) -> np.ndarray:
    return np.linspace(t_min, t_max, steps)          # float64, as the reference builds them


def threshold_sweep_counts(labels, probs, thresholds) -> np.ndarray:
    """int64 [steps][4] = (tp, fp, tn, fn) per threshold; the comparison happens in float64 like numpy's
    ``probs >= thresh`` on a float32 array and a float64 scalar."""
    labels = np.asarray(labels).astype(np.int64)
    p64 = np.asarray(probs).astype(np.float64)
    out = np.zeros((len(thresholds), 4), dtype=np.int64)
    for i, th in enumerate(thresholds):
        pred = p64 >= th
        pos = labels == 1
        out[i] = (np.sum(pred & pos), np.sum(pred & ~pos), np.sum(~pred & ~pos), np.sum(~pred & pos))
    return out


def metrics_from_counts(tp: int, fp: int, tn: int, fn: int):
    """accuracy, precision, recall, f1 exactly as sklearn computes them for average='binary', zero_division=0."""
    n = tp + fp + tn + fn
    acc = (tp + tn) / n if n else 0.0
    prec = tp / (tp + fp) if (tp + fp) else 0.0
    rec = tp / (tp + fn) if (tp + fn) else 0.0
    # sklearn (>= 1.3, pinned here: 1.9.0) forms F-beta from the counts, not from precision and recall
    f1 = (2 * tp / (2 * tp + fp + fn)) if (2 * tp + fp + fn) else 0.0
    return acc, prec, rec, f1


def find_optimal_threshold(labels, probs, t_min=0.3, t_max=0.7, steps=41):
    """(best_threshold, best_f1, best_acc, rows): train_advanced.py:243-265 (strict '>' keeps the first maximum)."""
    ths = thresholds_of(t_min, t_max, steps)
    counts = threshold_sweep_counts(labels, probs, ths)
    best_t, best_f1, best_acc = 0.5, 0, 0
    rows = []
    for th, (tp, fp, tn, fn) in zip(ths, counts):
        acc, prec, rec, f1 = metrics_from_counts(int(tp), int(fp), int(tn), int(fn))
        rows.append({"threshold": float(th), "accuracy": acc, "precision": prec, "recall": rec, "f1": f1})
        if f1 > best_f1:
            best_f1, best_t, best_acc = f1, float(th), acc
    return best_t, best_f1, best_acc, rows


def confusion_counts(y_true, y_pred):
    """(tn, fp, fn, tp) = sklearn.metrics.confusion_matrix(y_true, y_pred).ravel() for binary labels {0, 1}."""
    y_true = np.asarray(y_true).astype(np.int64)
    y_pred = np.asarray(y_pred).astype(np.int64)
    return (int(np.sum((y_true == 0) & (y_pred == 0))), int(np.sum((y_true == 0) & (y_pred == 1))),
            int(np.sum((y_true == 1) & (y_pred == 0))), int(np.sum((y_true == 1) & (y_pred == 1))))
