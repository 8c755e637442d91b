"""Generate tests/golden/threshold_golden.json by running the REAL reference functions in the build container:
find_optimal_threshold (/root/reference/train_advanced.py:239-275) and the confusion-matrix part of calculate_metrics
(/root/reference/test.py:241-243) on seeded synthetic scores.  TEST INFRASTRUCTURE ONLY; same stubbing of timm / wandb as
oracle/make_golden.py.  Usage: python oracle/make_golden_eval.py"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.make_golden import import_reference  # noqa: E402
from oracle import eval_oracle as eo  # noqa: E402


def make_case(seed, n, sharp, frac_live, dtype):
    rng = np.random.default_rng(seed)
    labels = (rng.random(n) < frac_live).astype(np.int64)
    logit = rng.normal(0.0, 1.0, n) + sharp * (labels * 2 - 1)
    probs = (1.0 / (1.0 + np.exp(-logit))).astype(dtype)
    if n > 8:   # scores sitting exactly on sweep thresholds (ties decide >= vs >)
        ths = np.linspace(0.3, 0.7, 41)
        probs[:4] = ths[[0, 10, 20, 40]].astype(dtype)
    return labels, probs


def main():
    ref = import_reference({"depth": 2})
    from sklearn.metrics import confusion_matrix

    class Cfg:
        threshold_min, threshold_max, threshold_steps = ref.Config.threshold_min, ref.Config.threshold_max, ref.Config.threshold_steps

    logged = []
    import wandb
    wandb.log = lambda d, *a, **k: logged.append(d)
    out = {"config": [Cfg.threshold_min, Cfg.threshold_max, Cfg.threshold_steps], "cases": []}
    for seed, n, sharp, frac, dt in [(1, 1747, 1.5, 0.7, np.float32), (2, 64, 0.3, 0.5, np.float32), (3, 5000, 3.0, 0.2, np.float16),
                                     (4, 7, 0.0, 0.5, np.float32), (5, 300, 1.0, 0.0, np.float32), (6, 300, 1.0, 1.0, np.float32)]:
        labels, probs = make_case(seed, n, sharp, frac, dt)
        logged.clear()
        best_t, best_f1, best_acc = ref.find_optimal_threshold(labels, probs, Cfg)
        rows = [{k.split("/")[1]: float(v) for k, v in d.items()} for d in logged]
        preds = (probs >= 0.5).astype(int)
        cm = confusion_matrix(labels, preds, labels=[0, 1]).ravel()
        out["cases"].append({"seed": seed, "n": n, "sharp": sharp, "frac_live": frac, "dtype": np.dtype(dt).name,
                             "best_threshold": float(best_t), "best_f1": float(best_f1), "best_acc": float(best_acc),
                             "rows": rows, "confusion_at_0.5": [int(v) for v in cm]})
        # the restatement must agree with the reference exactly
        t2, f2, a2, rows2 = eo.find_optimal_threshold(labels, probs, *out["config"])
        assert (t2, f2, a2) == (float(best_t), float(best_f1), float(best_acc)), (seed, t2, f2, a2, best_t, best_f1, best_acc)
        assert all(abs(r1[k] - r2[k]) == 0.0 for r1, r2 in zip(rows, rows2) for k in r2), seed
        assert list(eo.confusion_counts(labels, preds)) == [int(v) for v in cm]
        print("case", seed, "best", best_t, best_f1, best_acc)
    with open(os.path.join(ROOT, "tests", "golden", "threshold_golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote threshold_golden.json")


if __name__ == "__main__":
    main()
