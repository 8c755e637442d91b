"""Robustness sweep over batch sizes (ragged M = 197 B, partial CTA-pair tiles, fewer attention items than SMs ...):
the bf16 tensor-core path against the fp32 validation path of the same module (itself parity-tested against the oracle),
forward logits and a few gradients, plus batch invariance of sample 0."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vit_spoof_detection_pda_b200 as pkg  # noqa: E402
from oracle import vit_oracle as vo  # noqa: E402  (checker: seeded weights only)

dev = torch.device("cuda:0")
depth = 3
ref = vo.OracleViTFaceAntiSpoofing(dropout=0.0, depth=depth)
vo.seeded_init_(ref, seed=5)
models = {}
for prec in ("fp32", "bf16"):
    m = pkg.ViTFaceAntiSpoofing(dropout=0.0, depth=depth, precision=prec)
    m.load_state_dict(ref.state_dict())
    models[prec] = m.to(dev).train()
crit = pkg.FocalLoss(0.25, 2.0)
g = torch.Generator().manual_seed(3)
x_all = torch.randn(300, 3, 224, 224, generator=g).to(dev)
y_all = torch.randint(0, 2, (300,), generator=g).to(dev)
first = {}
worst = 0.0
for B in (1, 2, 5, 13, 17, 33, 64, 100, 129, 256, 300):
    x, y = x_all[:B], y_all[:B]
    outs, grads = {}, {}
    for prec, m in models.items():
        for p in m.parameters():
            p.grad = None
        out = m(x)
        crit(out, y).backward()
        outs[prec] = out.detach().float()
        grads[prec] = {n: p.grad.detach().float().clone() for n, p in m.named_parameters()
                       if n in ("vit.blocks.0.attn.qkv.weight", "vit.blocks.2.mlp.fc1.weight", "vit.patch_embed.proj.weight",
                                "classifier.2.weight", "vit.pos_embed", "vit.blocks.1.attn.qkv.bias")}
    err = float((outs["bf16"] - outs["fp32"]).abs().max())
    gerr = max(float((grads["bf16"][n] - grads["fp32"][n]).abs().max() / grads["fp32"][n].abs().max().clamp_min(1e-12)) for n in grads["fp32"])
    inv = 0.0
    for prec in outs:
        if prec in first:
            inv = max(inv, float((outs[prec][0] - first[prec]).abs().max()))
        else:
            first[prec] = outs[prec][0].clone()
    finite = all(torch.isfinite(v).all() for d in grads.values() for v in d.values())
    print(f"B={B:4d}  logits bf16-fp32 max-abs {err:.3e}   grads rel {gerr:.3e}   sample-0 drift vs B=1 {inv:.3e}  finite {finite}", flush=True)
    assert err < 2e-2 and gerr < 8e-2 and inv < 2e-2 and finite
    worst = max(worst, err)
print("batch sweep ok, worst logits err", worst)
