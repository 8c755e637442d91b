"""Host-side cost of enqueueing one training step (no GPU sync inside the loop): wall-clock per section."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vit_spoof_detection_pda_b200 as pkg  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = pkg.ViTFaceAntiSpoofing(dropout=0.1, depth=12, precision="bf16").to(dev).train()
crit = pkg.FocalLoss(0.25, 2.0)
opt = pkg.FusedAdam(model.parameters(), lr=1e-5, weight_decay=1e-4, adamw=False)
sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=1000, eta_min=1e-6)
x = torch.randn(64, 3, 224, 224, device=dev)
y = torch.randint(0, 2, (64,), device=dev)
acc = {}


def tick(name, t0):
    t = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (t - t0)
    return t


def step():
    t = time.perf_counter()
    out = model(x); t = tick("forward", t)
    loss, met = crit(out, y, with_metrics=True); t = tick("loss", t)
    loss.backward(); t = tick("backward", t)
    pkg.clip_grad_norm_(model.parameters(), 1.0); t = tick("clip", t)
    opt.step(); t = tick("adam", t)
    opt.zero_grad(set_to_none=True); t = tick("zero_grad", t)
    sched.step(); t = tick("sched", t)


for _ in range(5):
    step()
torch.cuda.synchronize()
acc.clear()
N = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
t_cpu = 0.0
for _ in range(N):
    torch.cuda.synchronize()      # empty launch queue: the sections below measure host work, not back-pressure
    t1 = time.perf_counter()
    step()
    t_cpu += time.perf_counter() - t1
e1.record()
torch.cuda.synchronize()
print(f"host enqueue {t_cpu / N * 1e3:.3f} ms/step (queue drained before every step), wall incl. drains {e0.elapsed_time(e1) / N:.3f} ms/step")
for k, v in acc.items():
    print(f"  {k:10s} {v / N * 1e3:7.3f} ms")

# ---- the same step captured into a CUDA graph (pkg.GraphedTrainStep): host cost of one replay
opt_g = pkg.FusedAdam(model.parameters(), lr=1e-5, weight_decay=1e-4, adamw=False, capturable=True)
gstep = pkg.GraphedTrainStep(model, crit, opt_g, x, y, max_grad_norm=1.0)
for _ in range(3):
    gstep(x, y)
torch.cuda.synchronize()
t_cpu = 0.0
e0.record()
for _ in range(N):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    gstep(x, y)
    t_cpu += time.perf_counter() - t1
e1.record()
torch.cuda.synchronize()
print(f"graphed step: host enqueue {t_cpu / N * 1e3:.3f} ms/step, wall incl. drains {e0.elapsed_time(e1) / N:.3f} ms/step")
e0.record()
for _ in range(N):
    gstep(x, y)
e1.record()
torch.cuda.synchronize()
print(f"graphed step back to back: {e0.elapsed_time(e1) / N:.3f} ms/step")
e0.record()
for _ in range(N):
    step()
e1.record()
torch.cuda.synchronize()
print(f"eager step back to back:   {e0.elapsed_time(e1) / N:.3f} ms/step")

if os.environ.get("PROFILE"):
    import cProfile
    import pstats
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        torch.cuda.synchronize()
        step()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
