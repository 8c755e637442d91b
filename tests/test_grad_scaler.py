"""the reference's fp16-style loop body (train_advanced.py:326-337: scale -> backward -> unscale_ -> clip ->
scaler.step -> scaler.update) driven through FusedGradScaler + FusedAdam must follow torch.amp.GradScaler + AdamW on the
oracle step for step -- including a step skipped because of a non-finite gradient, the scale backoff and the growth."""
import pytest
import torch

from oracle import vit_oracle as vo


def test_fused_grad_scaler_state_machine_without_gpu():
    """scale / backoff / growth bookkeeping against torch.amp.GradScaler driven through a real (tiny, CPU) optimizer."""
    from vit_spoof_detection_pda_b200.optim import FusedGradScaler
    s = FusedGradScaler(init_scale=8.0, growth_interval=2)
    t = torch.amp.GradScaler("cpu", init_scale=8.0, growth_interval=2)
    w = torch.nn.Parameter(torch.ones(3))
    opt = torch.optim.SGD([w], lr=0.1)
    for found_inf in [False, False, True, False, True, True, False, False, False]:
        t.scale((w * w).sum()).backward()
        if found_inf:
            w.grad[0] = float("nan")
        t.step(opt)
        t.update()
        opt.zero_grad(set_to_none=True)
        s._found_inf = found_inf
        s.update()
        assert s.get_scale() == t.get_scale()
        assert s.state_dict()["_growth_tracker"] == t.state_dict()["_growth_tracker"]
    sd = s.state_dict()
    s2 = FusedGradScaler()
    s2.load_state_dict(sd)
    assert s2.state_dict() == sd and set(sd) == set(t.state_dict())


# (lazy unscale, this package's clip, FusedAdam): every mix of fused and stock pieces the reference loop can be run with
# must give the reference's result (the lazy variants silently mis-clipped when mixed with stock pieces)
COMBOS = [("eager", "pkgclip", "fused"), ("lazy", "pkgclip", "fused"), ("eager", "stockclip", "fused"),
          ("eager", "pkgclip", "stockopt"), ("eager", "stockclip", "stockopt")]


@pytest.mark.gpu
@pytest.mark.parametrize("combo", COMBOS, ids=lambda c: "-".join(c))
def test_fused_grad_scaler_follows_torch_grad_scaler(combo):
    import vit_spoof_detection_pda_b200 as pkg
    unscale_mode, clip_mode, opt_mode = combo
    dev = torch.device("cuda:0")
    ref = vo.OracleViTFaceAntiSpoofing(dropout=0.0, depth=2)
    vo.seeded_init_(ref, seed=11)
    m = pkg.ViTFaceAntiSpoofing(dropout=0.0, depth=2, precision="fp32")
    m.load_state_dict(ref.state_dict())
    m = m.to(dev).train()
    ref.train()
    crit_ref, crit = vo.OracleFocalLoss(0.25, 2.0), pkg.FocalLoss(0.25, 2.0)
    opt_ref = torch.optim.AdamW(ref.parameters(), lr=3e-4, weight_decay=0.05)
    if opt_mode == "fused":
        opt = pkg.FusedAdam(m.parameters(), lr=3e-4, weight_decay=0.05, adamw=True)
    else:
        opt = torch.optim.AdamW(m.parameters(), lr=3e-4, weight_decay=0.05)
    clip = pkg.clip_grad_norm_ if clip_mode == "pkgclip" else torch.nn.utils.clip_grad_norm_
    sc_ref = torch.amp.GradScaler("cpu", init_scale=1024.0, growth_interval=2)
    sc = pkg.FusedGradScaler(init_scale=1024.0, growth_interval=2, lazy_unscale=(unscale_mode == "lazy"))
    for step in range(7):
        images, labels = vo.synthetic_batch(4, seed=100 + step)
        poison = step in (2, 3)
        # ---- reference loop body on the oracle (CPU, fp32)
        loss_r = crit_ref(ref(images), labels)
        sc_ref.scale(loss_r).backward()
        if poison:
            next(ref.parameters()).grad.view(-1)[0] = float("inf")
        sc_ref.unscale_(opt_ref)
        norm_r = torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        sc_ref.step(opt_ref)
        sc_ref.update()
        opt_ref.zero_grad(set_to_none=True)
        # ---- same body through the fused path
        loss = crit(m(images.to(dev)), labels.to(dev))
        sc.scale(loss).backward()
        if poison:
            next(m.parameters()).grad.view(-1)[0] = float("inf")
        sc.unscale_(opt)
        norm = clip(m.parameters(), 1.0)
        sc.step(opt)
        sc.update()
        opt.zero_grad(set_to_none=True)
        assert sc.get_scale() == sc_ref.get_scale(), step
        assert abs(float(loss) - float(loss_r)) <= 1e-4 * max(1.0, abs(float(loss_r))), step
        if not poison:
            assert abs(float(norm) - float(norm_r)) <= 1e-3 * float(norm_r), step
    assert sc.state_dict()["_growth_tracker"] == sc_ref.state_dict()["_growth_tracker"]
    for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        err = float((p.detach().cpu() - q.detach()).abs().max())
        # 5 applied AdamW steps of lr 3e-4 move a weight by up to 1.5e-3; Adam's sign-like update amplifies the fp32
        # summation-order noise of near-zero gradients, so the bound is relative to that total movement (3 %)
        assert err <= 5e-5, (combo, n, err)


@pytest.mark.gpu
def test_lazy_unscale_rejects_stock_optimizer():
    import vit_spoof_detection_pda_b200 as pkg
    m = pkg.ViTFaceAntiSpoofing(dropout=0.0, depth=1, precision="fp32").to("cuda:0").train()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
    sc = pkg.FusedGradScaler(lazy_unscale=True)
    images, labels = vo.synthetic_batch(2, seed=1)
    sc.scale(pkg.FocalLoss()(m(images.cuda()), labels.cuda())).backward()
    with pytest.raises(TypeError):
        sc.unscale_(opt)


@pytest.mark.gpu
def test_pkg_clip_scales_in_place_without_fused_adam():
    """clip_grad_norm_ of this package in front of a stock torch optimizer must clip right away (there is no fused pass
    to fold it into)."""
    import vit_spoof_detection_pda_b200 as pkg
    m = pkg.ViTFaceAntiSpoofing(dropout=0.0, depth=1, precision="fp32").to("cuda:0").train()
    images, labels = vo.synthetic_batch(4, seed=3)
    (pkg.FocalLoss()(m(images.cuda()), labels.cuda()) * 1000.0).backward()
    before = torch.cat([p.grad.reshape(-1) for p in m.parameters()]).norm()
    ret = pkg.clip_grad_norm_(m.parameters(), 0.5)
    after = torch.cat([p.grad.reshape(-1) for p in m.parameters()]).norm()
    assert float(before) > 0.5
    assert abs(float(ret) - float(before)) <= 1e-4 * float(before)
    assert abs(float(after) - 0.5) <= 1e-4
