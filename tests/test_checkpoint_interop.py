"""the checkpoint the reference writes (train_advanced.py:475-489: model / optimizer / scheduler / scaler
state dicts) is interchangeable between the fused objects and the stock torch ones: FusedAdam.state_dict() loads into
torch.optim.AdamW on the oracle model (same per-parameter exp_avg / exp_avg_sq / step layout) and back, and training
continues identically from either side."""
import io

import pytest
import torch

from oracle import vit_oracle as vo


@pytest.mark.gpu
def test_checkpoint_round_trip_between_fused_and_torch_optimizers():
    import vit_spoof_detection_pda_b200 as pkg
    dev = torch.device("cuda:0")
    ref = vo.OracleViTFaceAntiSpoofing(dropout=0.0, depth=2)
    vo.seeded_init_(ref, seed=21)
    m = pkg.ViTFaceAntiSpoofing(dropout=0.0, depth=2, precision="fp32")
    m.load_state_dict(ref.state_dict())
    m = m.to(dev).train()
    crit = pkg.FocalLoss(0.25, 2.0)
    opt = pkg.FusedAdam(m.parameters(), lr=3e-4, weight_decay=0.05, adamw=True)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=20, eta_min=1e-6)
    scaler = pkg.FusedGradScaler(init_scale=256.0)

    def fused_step(seed):
        images, labels = vo.synthetic_batch(4, seed=seed)
        loss = crit(m(images.to(dev)), labels.to(dev))
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        pkg.clip_grad_norm_(m.parameters(), 1.0)
        scaler.step(opt)
        scaler.update()
        opt.zero_grad(set_to_none=True)
        sched.step()

    for s in range(3):
        fused_step(300 + s)
    # ---- the reference's checkpoint container, through torch.save / torch.load(weights_only=False) (test.py:174)
    ckpt = {"epoch": 1, "model_state_dict": m.state_dict(), "optimizer_state_dict": opt.state_dict(),
            "scheduler_state_dict": sched.state_dict(), "scaler_state_dict": scaler.state_dict(), "metrics": {"f1": 0.5},
            "config": {"learning_rate": 3e-4}}
    buf = io.BytesIO()
    torch.save(ckpt, buf)
    buf.seek(0)
    ck = torch.load(buf, weights_only=False, map_location="cpu")

    # ---- resume on the STOCK side: oracle model + torch.optim.AdamW + torch GradScaler
    ref.load_state_dict(ck["model_state_dict"])
    ref.train()
    opt_t = torch.optim.AdamW(ref.parameters(), lr=3e-4, weight_decay=0.05)
    opt_t.load_state_dict(ck["optimizer_state_dict"])
    sched_t = torch.optim.lr_scheduler.CosineAnnealingLR(opt_t, T_max=20, eta_min=1e-6)
    sched_t.load_state_dict(ck["scheduler_state_dict"])
    sc_t = torch.amp.GradScaler("cpu")
    sc_t.load_state_dict(ck["scaler_state_dict"])
    crit_t = vo.OracleFocalLoss(0.25, 2.0)
    assert all(int(st["step"]) == 3 for st in opt_t.state_dict()["state"].values())
    for s in range(2):
        images, labels = vo.synthetic_batch(4, seed=310 + s)
        loss = crit_t(ref(images), labels)
        sc_t.scale(loss).backward()
        sc_t.unscale_(opt_t)
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        sc_t.step(opt_t)
        sc_t.update()
        opt_t.zero_grad(set_to_none=True)
        sched_t.step()
        fused_step(310 + s)       # the fused side continues in lock step
    for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        assert float((p.detach().cpu() - q.detach()).abs().max()) <= 5e-5, n
    assert opt.param_groups[0]["lr"] == pytest.approx(opt_t.param_groups[0]["lr"], rel=1e-12)

    # ---- and back: the stock optimizer's state loads into a fresh fused optimizer
    m2 = pkg.ViTFaceAntiSpoofing(dropout=0.0, depth=2, precision="fp32")
    m2.load_state_dict(ref.state_dict())
    m2 = m2.to(dev).train()
    opt2 = pkg.FusedAdam(m2.parameters(), lr=3e-4, weight_decay=0.05, adamw=True)
    opt2.load_state_dict(opt_t.state_dict())
    sd_a, sd_b = opt2.state_dict(), opt_t.state_dict()
    assert set(sd_a["state"]) == set(sd_b["state"])
    for k in sd_b["state"]:
        assert float(sd_a["state"][k]["step"]) == float(sd_b["state"][k]["step"])
        assert torch.equal(sd_a["state"][k]["exp_avg"].cpu(), sd_b["state"][k]["exp_avg"])
        assert torch.equal(sd_a["state"][k]["exp_avg_sq"].cpu(), sd_b["state"][k]["exp_avg_sq"])
