"""Generate tests/golden/* by running the REAL reference classes in the build container.

TEST INFRASTRUCTURE ONLY (see oracle/vit_oracle.py header).  Runs only where /root/reference
exists (the build container); the GPU box consumes the committed fixtures.

How the reference is executed: /root/reference/train_advanced.py imports ``timm`` (absent, not
installable) and ``wandb`` (needs a service).  Both are replaced by stubs in ``sys.modules``
BEFORE the import: ``timm.create_model`` returns ``OracleViTEncoder`` (the restated third-party
arithmetic), ``wandb.log`` is a no-op.  Everything else -- ``FocalLoss`` (train_advanced.py:90-107),
``ViTFaceAntiSpoofing`` (187-204), ``train_epoch`` (315-365), ``validate``'s loop body --
is the reference's own unmodified code.

Fixtures written (small; the 345 MB of weights are re-derived from a seed, never stored):
  focal_golden.pt       reference FocalLoss values + autograd dlogits over the sweep grid
  model_golden.pt       reference ViTFaceAntiSpoofing logits / loss / per-tensor grad summaries
                        (norm + first 4 elements of every one of the 156 grads), depth 12 and depth 2
  train_golden.pt       reference train_epoch over 4 batches of 2 (AdamW code config and
                        Adam README config): per-step loss, final parameter checksums
  encoder_xcheck.json   OracleViTEncoder vs torchvision.models.vit_b_16 max-abs feature diff

Usage:  python oracle/make_golden.py            (about 1-2 min on 8 cores)
"""
from __future__ import annotations

import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

from oracle import vit_oracle as vo  # noqa: E402


def import_reference(depth_holder):
    timm = types.ModuleType("timm")

    def create_model(name, pretrained=False, num_classes=0, **kw):
        assert name == "vit_base_patch16_224" and num_classes == 0
        return vo.OracleViTEncoder(depth=depth_holder["depth"])

    timm.create_model = create_model
    sys.modules["timm"] = timm
    wandb = types.ModuleType("wandb")
    wandb.log = lambda *a, **k: None
    wandb.init = lambda *a, **k: None
    wandb.config = {}
    sys.modules["wandb"] = wandb
    sys.path.insert(0, REF)
    import train_advanced as ref  # the reference module itself
    return ref


def grad_summary(model):
    out = {}
    for n, p in model.named_parameters():
        g = p.grad.detach().flatten()
        out[n] = {"norm": float(g.double().norm()), "head": g[:4].clone(), "absmax": float(g.abs().max())}
    return out


def main():
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    os.makedirs(OUT, exist_ok=True)
    depth_holder = {"depth": 12}
    ref = import_reference(depth_holder)

    class Cfg:
        model_name = "vit_base_patch16_224"
        pretrained = False
        num_classes = 2
        dropout = 0.0
        device = "cpu"
        mixed_precision = False
        gradient_accumulation_steps = 1
        max_grad_norm = 1.0
        num_epochs = 1
        log_interval = 10

    # ---------------- focal loss: the real class over the sweep grid ----------------
    g = torch.Generator().manual_seed(7)
    logits = torch.randn(16, 2, generator=g) * 3
    logits[0] = torch.tensor([30.0, -30.0])     # pt -> 1 or 0 depending on the label
    logits[1] = torch.tensor([-30.0, 30.0])
    logits[2] = torch.tensor([0.0, 0.0])
    logits[3] = torch.tensor([88.0, -88.0])
    targets = torch.randint(0, 2, (16,), generator=g)
    targets[0], targets[1], targets[3] = 0, 0, 1
    focal = {"logits": logits, "targets": targets, "cases": []}
    for gamma in (1.5, 2.0, 2.5):
        for alpha in (0.15, 0.25, 0.35):
            for red in ("mean", "sum"):
                z = logits.clone().requires_grad_(True)
                loss = ref.FocalLoss(alpha=alpha, gamma=gamma, reduction=red)(z, targets)
                loss.backward()
                focal["cases"].append({"alpha": alpha, "gamma": gamma, "reduction": red,
                                       "loss": loss.detach().clone(), "dlogits": z.grad.clone()})
    z = logits.clone()
    focal["none_reduction"] = ref.FocalLoss(alpha=0.25, gamma=2.0, reduction="none")(z, targets).clone()
    torch.save(focal, os.path.join(OUT, "focal_golden.pt"))
    print("focal cases:", len(focal["cases"]))

    # ---------------- model wrapper + head + loss: forward / backward ----------------
    model_g = {}
    for depth, batch in ((12, 2), (2, 3)):
        depth_holder["depth"] = depth
        m = ref.ViTFaceAntiSpoofing(Cfg)
        vo.seeded_init_(m, seed=42)
        names = [n for n, _ in m.state_dict().items()]
        spec = vo.expected_state_dict_spec(depth)
        assert names == [n for n, _ in spec], "state_dict key contract broken"
        images, labels = vo.synthetic_batch(batch, seed=42)
        m.train()
        out = m(images)
        loss = ref.FocalLoss(alpha=Cfg.__dict__.get("focal_alpha", 0.25), gamma=2.0)(out, labels)
        loss.backward()
        model_g[f"depth{depth}"] = {"batch": batch, "seed": 42, "logits": out.detach().clone(),
                                    "loss": loss.detach().clone(), "grads": grad_summary(m),
                                    "labels": labels.clone()}
        print(f"depth {depth}: logits {out.detach().flatten().tolist()} loss {loss.item():.6f}")
    torch.save(model_g, os.path.join(OUT, "model_golden.pt"))

    # ---------------- reference train_epoch over 4 tiny batches ----------------
    train_g = {}
    depth_holder["depth"] = 2
    for kind, lr, wd in (("adamw", 3e-4, 0.05), ("adam", 1e-5, 1e-4)):
        m = ref.ViTFaceAntiSpoofing(Cfg)
        vo.seeded_init_(m, seed=42)
        loader = [vo.synthetic_batch(2, seed=100 + i) for i in range(4)]
        opt = vo.make_optimizer(m.parameters(), kind, lr=lr, weight_decay=wd)
        sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=4, eta_min=1e-6)
        scaler = ref.GradScaler(enabled=False)
        crit = ref.FocalLoss(alpha=0.25, gamma=2.0)
        losses = []
        orig_update = ref.AverageMeter.update

        def spy(self, val, n=1, _o=orig_update):
            losses.append(val)
            return _o(self, val, n)

        ref.AverageMeter.update = spy
        try:
            ref.train_epoch(m, loader, crit, opt, sched, scaler, Cfg, 0, 0)
        finally:
            ref.AverageMeter.update = orig_update
        train_g[kind] = {"lr": lr, "wd": wd, "loss_per_step": losses[0::2], "acc_per_step": losses[1::2],
                         "param_norms": {n: float(p.detach().double().norm()) for n, p in m.named_parameters()},
                         "param_heads": {n: p.detach().flatten()[:4].clone() for n, p in m.named_parameters()}}
        print(kind, "losses", losses[0::2])
    torch.save(train_g, os.path.join(OUT, "train_golden.pt"))

    # ---------------- encoder cross-check vs torchvision ----------------
    import torchvision
    enc = vo.OracleViTEncoder(12)
    vo.seeded_init_(enc, seed=42)
    tv = torchvision.models.vit_b_16(weights=None)
    tv.heads = torch.nn.Identity()
    missing = tv.load_state_dict(vo.encoder_state_to_torchvision(enc.state_dict()), strict=True)
    images, _ = vo.synthetic_batch(2, seed=42)
    with torch.no_grad():
        a = enc.eval()(images)
        b = tv.eval()(images)
    diff = float((a - b).abs().max())
    rel = diff / float(b.abs().max())
    with open(os.path.join(OUT, "encoder_xcheck.json"), "w") as f:
        json.dump({"against": f"torchvision {torchvision.__version__} vit_b_16", "torch": torch.__version__,
                   "batch": 2, "max_abs_diff": diff, "rel_to_absmax": rel, "strict_load": str(missing)}, f, indent=1)
    print("torchvision cross-check max-abs diff", diff, "rel", rel)
    assert rel < 1e-4


if __name__ == "__main__":
    main()
