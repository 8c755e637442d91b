"""device-side threshold sweep + confusion counts, and the uint8 HWC input edge.

CPU part: the numpy restatement (oracle/eval_oracle.py) reproduces the golden outputs of the REAL reference functions
(tests/golden/threshold_golden.json, written by oracle/make_golden_eval.py from train_advanced.find_optimal_threshold and
sklearn's confusion_matrix).  GPU part: the CUDA kernels reproduce them bit for bit (integer counts, hence identical
floats), including scores exactly on a threshold, fp16 scores, all-live / all-spoof label sets and streaming updates."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import eval_oracle as eo
from oracle.make_golden_eval import make_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "threshold_golden.json")))
DT = {"float32": np.float32, "float16": np.float16}


def _inputs(case):
    return make_case(case["seed"], case["n"], case["sharp"], case["frac_live"], DT[case["dtype"]])


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"seed{c['seed']}-n{c['n']}")
def test_oracle_matches_reference_golden(case):
    labels, probs = _inputs(case)
    t, f1, acc, rows = eo.find_optimal_threshold(labels, probs, *GOLD["config"])
    assert (t, f1, acc) == (case["best_threshold"], case["best_f1"], case["best_acc"])
    for r, g in zip(rows, case["rows"]):
        assert r == g                       # exact float equality, every threshold
    assert list(eo.confusion_counts(labels, (probs >= 0.5).astype(int))) == case["confusion_at_0.5"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: f"seed{c['seed']}-n{c['n']}")
def test_device_threshold_sweep_bit_exact(case):
    import vit_spoof_detection_pda_b200 as pkg
    labels, probs = _inputs(case)
    dev = torch.device("cuda:0")

    class Cfg:
        threshold_min, threshold_max, threshold_steps = GOLD["config"]

    p = torch.from_numpy(probs).to(dev)
    y = torch.from_numpy(labels).to(dev)
    sweep = pkg.ThresholdSweep(Cfg, device=dev)
    # streamed in uneven batches, as a validation loop would
    for lo in range(0, len(labels), 257):
        sweep.update(p[lo:lo + 257], y[lo:lo + 257])
    counts = sweep.counts().cpu().numpy()
    ref_counts = eo.threshold_sweep_counts(labels, probs, eo.thresholds_of(*GOLD["config"]))
    assert np.array_equal(counts, ref_counts)
    assert sweep.results() == case["rows"]
    assert sweep.best() == (case["best_threshold"], case["best_f1"], case["best_acc"])
    # the drop-in function (reference signature) on device tensors and on host arrays
    assert pkg.find_optimal_threshold(y, p, Cfg) == (case["best_threshold"], case["best_f1"], case["best_acc"])
    assert pkg.find_optimal_threshold(labels, probs, Cfg) == (case["best_threshold"], case["best_f1"], case["best_acc"])
    preds = (p >= 0.5).to(torch.int64)
    assert list(pkg.confusion_counts(y, preds)) == case["confusion_at_0.5"]


@pytest.mark.gpu
def test_device_threshold_sweep_nan_and_empty():
    import vit_spoof_detection_pda_b200 as pkg
    dev = torch.device("cuda:0")
    sweep = pkg.ThresholdSweep(threshold_min=0.3, threshold_max=0.7, threshold_steps=41, device=dev)
    sweep.update(torch.empty(0, device=dev), torch.empty(0, dtype=torch.int64, device=dev))
    assert int(sweep.counts().sum()) == 0 and sweep.best() == (0.5, 0, 0)
    p = torch.tensor([float("nan"), 0.9, 0.1], device=dev)
    y = torch.tensor([1, 1, 0], device=dev)
    sweep.update(p, y)
    ref = eo.threshold_sweep_counts(y.cpu().numpy(), p.cpu().numpy(), sweep.thresholds)
    assert np.array_equal(sweep.counts().cpu().numpy(), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_uint8_hwc_input_edge_equals_normalised_float_input(precision):
    """n2: feeding raw uint8 HWC pixels gives exactly the logits of feeding ToTensor+Normalize'd fp32 NCHW images
    (train_advanced.py:180-181 / test.py get_test_transforms), because the fused loader performs the same fp32 operations in
    the same order before the same cast."""
    import vit_spoof_detection_pda_b200 as pkg
    from oracle import vit_oracle as vo
    dev = torch.device("cuda:0")
    ref = vo.OracleViTFaceAntiSpoofing(dropout=0.0, depth=2)
    vo.seeded_init_(ref, seed=7)
    m = pkg.ViTFaceAntiSpoofing(dropout=0.0, depth=2, precision=precision)
    m.load_state_dict(ref.state_dict())
    m = m.to(dev).eval()
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (3, 224, 224, 3), dtype=torch.uint8, generator=g)
    mean = torch.tensor(m.pixel_mean, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(m.pixel_std, dtype=torch.float32).view(1, 3, 1, 1)
    x = (u8.permute(0, 3, 1, 2).to(torch.float32).div(255.0) - mean) / std      # ToTensor + Normalize
    with torch.no_grad():
        a = m(u8.to(dev))
        b = m(x.to(dev))
        o = ref.eval()(x)
    assert torch.equal(a, b)
    tol = 1e-4 if precision == "fp32" else 2e-2
    assert float((a.cpu() - o).abs().max()) <= tol * max(1.0, float(o.abs().max()))


@pytest.mark.gpu
def test_host_scalars_deliver_every_value_one_step_late():
    """pkg.HostScalars (the asynchronous stand-in for loss.item() / acc.item(), train_advanced.py:345-346) hands over
    exactly the pushed values, in order, one push late, and flush() returns the last ones."""
    import vit_spoof_detection_pda_b200 as pkg
    dev = torch.device("cuda:0")
    reader = pkg.HostScalars(dev, slots=3)
    got = []
    vals = [(float(i) * 0.25 + 1e-3, float(7 * i)) for i in range(8)]
    for a, b in vals:
        ta, tb = torch.tensor(a, device=dev, dtype=torch.float32), torch.tensor(int(b), device=dev, dtype=torch.int32)
        prev = reader.push(ta, tb)
        if prev is not None:
            got.append(prev)
    got.append(reader.flush())
    assert reader.flush() is None
    assert len(got) == len(vals)
    for (a, b), (ga, gb) in zip(vals, got):
        assert ga == float(torch.tensor(a, dtype=torch.float32)) and gb == b
