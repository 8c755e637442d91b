"""The practical bar on the same GPU ("additional recommended baseline"): the training step of
train_advanced.py:322-346 written with stock PyTorch library code only -- torchvision's `vit_b_16` encoder (same
architecture as timm's vit_base_patch16_224, see oracle header for the key remap), bf16 autocast, SDPA inside
nn.MultiheadAttention, focal loss from ATen ops, `clip_grad_norm_`, fused/foreach AdamW.  No code from this package
and nothing from oracle/ is imported: the number is what a user of the reference gets by moving the unmodified
script to a B200.

    python tools/torch_eager_baseline.py [--batch 64] [--steps 20] [--warmup 5]
prints one JSON line {"impl": "torch_eager", "img_s": ..., "ms_per_step": ..., "bs1_latency_ms_p50": ...}.
"""
import argparse
import json

import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision


class EagerPAD(nn.Module):
    def __init__(self, dropout=0.1):
        super().__init__()
        self.vit = torchvision.models.vit_b_16(weights=None)
        self.vit.heads = nn.Identity()
        self.classifier = nn.Sequential(nn.LayerNorm(768), nn.Dropout(dropout), nn.Linear(768, 512), nn.GELU(),
                                        nn.Dropout(dropout), nn.Linear(512, 2))

    def forward(self, x):
        return self.classifier(self.vit(x))


def focal(logits, y, alpha=0.25, gamma=2.0):
    ce = F.cross_entropy(logits, y, reduction="none")
    pt = torch.exp(-ce)
    return (alpha * (1 - pt) ** gamma * ce).mean()


def run(batch=64, steps=20, warmup=5, with_inference=True, device="cuda:0"):
    dev = torch.device(device)
    torch.manual_seed(42)
    torch.backends.cuda.matmul.allow_tf32 = True
    model = EagerPAD().to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-5, weight_decay=1e-4, fused=True)
    x = torch.randn(batch, 3, 224, 224, device=dev)
    y = torch.randint(0, 2, (batch,), device=dev)

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(x)
            loss = focal(out.float(), y)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    res = {"impl": "torch_eager", "torch": torch.__version__, "batch": batch, "steps": steps,
           "ms_per_step": ms, "img_s": batch / (ms * 1e-3), "tflops": batch / (ms * 1e-3) * 105.150e9 / 1e12,
           "what": "torchvision vit_b_16 + the reference head, bf16 autocast, SDPA, clip_grad_norm_, fused AdamW (library kernels only)"}
    if not with_inference:
        return res

    # batch-1 eval latency, eager launches (what test.py does) and CUDA-graph replay
    model.eval()
    x1 = torch.randn(1, 3, 224, 224, device=dev)
    lat = []
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        for _ in range(20):
            model(x1)
        torch.cuda.synchronize()
        for _ in range(200):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            model(x1)
            e.record()
            e.synchronize()
            lat.append(s.elapsed_time(e))
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            model(x1)
            with torch.cuda.graph(g, stream=side):
                model(x1)
        torch.cuda.synchronize()
        glat = []
        for _ in range(300):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            g.replay()
            e.record()
            e.synchronize()
            glat.append(s.elapsed_time(e))
        xb = torch.randn(256, 3, 224, 224, device=dev)
        for _ in range(3):
            model(xb)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            model(xb)
        e1.record()
        torch.cuda.synchronize()
        infer = 256 * 10 / (e0.elapsed_time(e1) * 1e-3)
    lat.sort()
    glat.sort()
    res.update({"bs1_latency_ms_p50_eager": lat[len(lat) // 2], "bs1_latency_ms_p50_graph": glat[len(glat) // 2],
                "bs256_infer_img_s": infer})
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    a = ap.parse_args()
    print(json.dumps(run(a.batch, a.steps, a.warmup)))


if __name__ == "__main__":
    main()
