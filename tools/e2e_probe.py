"""Where does the end-to-end loop (pinned host batches + two .item() syncs per step) lose time against the
device-resident loop?  Times four variants of the same 20 steps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vit_spoof_detection_pda_b200 as pkg  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = pkg.ViTFaceAntiSpoofing(dropout=0.1, depth=12, precision="bf16").to(dev).train()
crit = pkg.FocalLoss(0.25, 2.0)
opt = pkg.FusedAdam(model.parameters(), lr=1e-5, weight_decay=1e-4, adamw=False)
B, N = 64, 20
host = [(torch.randn(B, 3, 224, 224).pin_memory(), torch.randint(0, 2, (B,)).pin_memory()) for _ in range(4)]
devb = [(a.to(dev), b.to(dev)) for a, b in host]


def step(x, y):
    loss, met = crit(model(x), y, with_metrics=True)
    loss.backward()
    pkg.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()
    opt.zero_grad(set_to_none=True)
    return loss, met


reader = pkg.HostScalars(dev)


def run(name, prefetch, sync):
    def batches():
        for i in range(N):
            yield host[i % 4] if prefetch else devb[i % 4]
    it = pkg.DevicePrefetcher(batches(), dev) if prefetch else batches()
    for _ in range(3):
        step(*devb[0])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for x, y in it:
        loss, met = step(x, y)
        if sync == 3:
            reader.push(loss, met["ncorrect"])
        elif sync == 2:
            loss.item(); met["ncorrect"].item()
        elif sync == 1:
            loss.item()
    if sync == 3:
        reader.flush()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:46s} {e0.elapsed_time(e1) / N:7.3f} ms/step", flush=True)


run("device-resident, no sync", False, 0)
run("device-resident, loss.item()", False, 1)
run("device-resident, loss.item() + ncorrect.item()", False, 2)
run("pinned host batches (prefetcher), no sync", True, 0)
run("pinned host batches + both .item()", True, 2)
run("device-resident + HostScalars", False, 3)
run("pinned host batches + HostScalars (bench e2e)", True, 3)
run("device-resident, no sync (again)", False, 0)

