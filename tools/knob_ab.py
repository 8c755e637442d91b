"""Same-process A/B of vitk_debug_set knobs on the bench workload (bs 64, bf16 training step): alternates the given knob
settings over several rounds of timed steps, so box-to-box and thermal drift cancel.
    python tools/knob_ab.py 12:0 12:1 [--rounds 4] [--steps 10]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vit_spoof_detection_pda_b200 as pkg  # noqa: E402
from vit_spoof_detection_pda_b200 import _lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("settings", nargs="+", help="comma-separated key:value lists, one per arm, e.g. 12:0 12:1")
ap.add_argument("--rounds", type=int, default=4)
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
lib = L.load()
model = pkg.ViTFaceAntiSpoofing(dropout=0.1, depth=12, precision="bf16").to(dev).train()
crit = pkg.FocalLoss(0.25, 2.0)
opt = pkg.FusedAdam(model.parameters(), lr=1e-5, weight_decay=1e-4, adamw=False)
xs = [torch.randn(64, 3, 224, 224, device=dev) for _ in range(4)]
ys = [torch.randint(0, 2, (64,), device=dev) for _ in range(4)]


def step(i):
    loss, _ = crit(model(xs[i % 4]), ys[i % 4], with_metrics=True)
    loss.backward()
    pkg.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()
    opt.zero_grad(set_to_none=True)


def apply(setting):
    for kv in setting.split(","):
        k, v = kv.split(":")
        if k == "noprezero":       # host-side switch: gradient-buffer clear at the head of backward instead of under the forward
            os.environ["VITK_NO_PREZERO"] = v
        else:
            lib.vitk_debug_set(int(k), int(v))


for i in range(5):
    step(i)
res = {s: [] for s in a.settings}
for r in range(a.rounds):
    for s in a.settings:
        apply(s)
        for i in range(2):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        res[s].append(e0.elapsed_time(e1) / a.steps)
for s, v in res.items():
    print(f"knobs {s:12s} ms/step: " + " ".join(f"{t:.3f}" for t in v) + f"   mean {sum(v) / len(v):.3f}")
