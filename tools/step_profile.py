"""Runs 3 training steps of the bench workload (bs 64, bf16, eager launches) -- the command ncu wraps for the launch list.
The third step is bracketed by cudaProfilerStart/Stop: `ncu --profile-from-start off ...` captures exactly one whole step
(no -s / -c arithmetic); without that flag every launch is captured as before."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vit_spoof_detection_pda_b200 as pkg  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
B = int(os.environ.get("VITK_PROFILE_BATCH", "64"))
model = pkg.ViTFaceAntiSpoofing(dropout=0.1, depth=12, precision="bf16").to(dev).train()
crit = pkg.FocalLoss(0.25, 2.0)
opt = pkg.FusedAdam(model.parameters(), lr=1e-5, weight_decay=1e-4, adamw=False)
x = torch.randn(B, 3, 224, 224, device=dev)
y = torch.randint(0, 2, (B,), device=dev)
for step in range(3):
    if step == 2:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    loss, _ = crit(model(x), y, with_metrics=True)
    loss.backward()
    pkg.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()
    opt.zero_grad(set_to_none=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", loss.item())
