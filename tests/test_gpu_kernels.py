"""Per-kernel parity of the C-ABI entry points against plain torch fp32 references on the GPU box.
Tolerances: fp32-validate kernels 1e-4 relative (north_star); bf16 kernels 2e-2 max-abs on O(1) data
(relative to the output scale)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from vit_spoof_detection_pda_b200 import _lib as L
    import kernels_api as K
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    DEV = torch.device("cuda:0")
else:  # collected but skipped on the CPU box
    L = K = DEV = None

FP32_TOL = 1e-4
BF16_TOL = 2e-2


def _g(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


def randn(*shape, seed=0, scale=1.0, dtype=torch.float32):
    return (torch.randn(*shape, generator=_g(seed), device=DEV) * scale).to(dtype)


# ------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("rows", [1, 7, 197, 1576])
@pytest.mark.parametrize("out_dtype", ["f32", "bf16"])
def test_layernorm_fwd_bwd(rows, out_dtype):
    odt = torch.float32 if out_dtype == "f32" else torch.bfloat16
    x = randn(rows, 768, seed=1, scale=2.0) + 0.5
    gamma = 1 + 0.1 * randn(768, seed=2)
    beta = 0.1 * randn(768, seed=3)
    y, mean, rstd = K.layernorm_fwd(x, gamma, beta, 1e-6, odt)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (768,), gr, br, 1e-6)
    tol = FP32_TOL if odt == torch.float32 else BF16_TOL
    assert K.rel_err(y.float(), yr) < tol
    dy = randn(rows, 768, seed=4).to(odt)
    dres = randn(rows, 768, seed=5)
    yr.backward(dy.float())
    dx, dx16, dg, db, cs = K.layernorm_bwd(dy, x, gamma, mean, rstd, dres=dres.clone(), want16=True)
    assert K.rel_err(dx, xr.grad + dres) < FP32_TOL
    assert K.rel_err(cs, (xr.grad + dres).sum(0)) < FP32_TOL * 5
    assert K.rel_err(dx16.float(), xr.grad + dres) < BF16_TOL
    assert K.rel_err(dg, gr.grad) < FP32_TOL * 5
    assert K.rel_err(db, br.grad) < FP32_TOL * 5


def test_layernorm_strided_cls_rows():
    B = 5
    x = randn(B * 197, 768, seed=7)
    gamma, beta = 1 + 0.1 * randn(768, seed=8), 0.1 * randn(768, seed=9)
    y, _, _ = K.layernorm_fwd(x, gamma, beta, 1e-6, torch.float32, x_stride=197 * 768, rows=B)
    ref = F.layer_norm(x.view(B, 197, 768)[:, 0], (768,), gamma, beta, 1e-6)
    assert K.rel_err(y, ref) < FP32_TOL


# ------------------------------------------------------------------ Linear (all engines)
ENGINE_CASES = [("f32", "simt"), ("bf16", "simt"), ("bf16", "tcgen05")]


def _case(case):
    dtype = torch.float32 if case[0] == "f32" else torch.bfloat16
    return dtype, [L.ENGINE_SIMT if case[1] == "simt" else L.ENGINE_TCGEN05]


def _tol(dtype):
    return FP32_TOL if dtype == torch.float32 else BF16_TOL


SHAPES = [(197, 768, 768), (394, 2304, 768), (1576, 3072, 768), (1000, 768, 3072)]


@pytest.mark.parametrize("case", ENGINE_CASES, ids=lambda c: f"{c[0]}-{c[1]}")
@pytest.mark.parametrize("M,N,Kd", SHAPES)
def test_linear_fwd_epilogues(case, M, N, Kd):
    dtype, engines = _case(case)
    x = randn(M, Kd, seed=11).to(dtype)
    w = randn(N, Kd, seed=12, scale=0.05).to(dtype)
    b = randn(N, seed=13, scale=0.5)
    res = randn(M, N, seed=14)
    ref = x.float() @ w.float().t() + b
    for eng in engines:
        y = K.linear_fwd(x, w, b, L.EPI_BIAS, eng)
        assert K.rel_err(y.float(), ref) < _tol(dtype), ("bias", eng)
        g, dg = K.linear_fwd(x, w, b, L.EPI_BIAS_GELU, eng)
        ur = ref.to(dtype).float().requires_grad_(True)      # GELU acts on the activation-dtype fc1 output
        gr = F.gelu(ur)
        gr.sum().backward()
        assert K.rel_err(g.float(), gr) < _tol(dtype), ("gelu-g", eng)
        assert K.rel_err(dg.float(), ur.grad) < _tol(dtype), ("gelu-dg", eng)
        y = K.linear_fwd(x, w, b, L.EPI_BIAS_RESIDUAL, eng, residual=res)
        assert K.rel_err(y, ref + res) < _tol(dtype), ("residual", eng)
        y = K.linear_fwd(x, w, b, L.EPI_QKV_SCATTER, eng)
        assert K.rel_err(K.from_headmajor(y).float(), ref) < _tol(dtype), ("scatter", eng)


@pytest.mark.parametrize("case", ENGINE_CASES, ids=lambda c: f"{c[0]}-{c[1]}")
@pytest.mark.parametrize("M,N,Kd", SHAPES)
def test_linear_dgrad(case, M, N, Kd):
    dtype, engines = _case(case)
    dy = randn(M, N, seed=21).to(dtype)
    w = randn(N, Kd, seed=22, scale=0.05).to(dtype)
    u = randn(M, Kd, seed=23).to(dtype)
    ref = dy.float() @ w.float()
    gp = torch.autograd.functional.jvp  # noqa: F841 (documentation: gelu' from autograd below)
    uu = u.float().clone().requires_grad_(True)
    F.gelu(uu).sum().backward()
    for eng in engines:
        dx = K.linear_dgrad(dy, w, eng)
        assert K.rel_err(dx.float(), ref) < _tol(dtype), ("plain", eng)
        dx, cs = K.linear_dgrad(dy, w, eng, want_colsum=True)
        assert K.rel_err(dx.float(), ref) < _tol(dtype), ("plain + column sums", eng)
        assert K.rel_err(cs, dx.float().sum(0)) < 1e-3, ("plain: fused column sums", eng)
        dx, cs = K.linear_dgrad(dy, w, eng, gelu_grad=u, want_colsum=True)
        assert K.rel_err(dx.float(), ref * u.float()) < _tol(dtype), ("gelu_bwd", eng)
        assert K.rel_err(cs, dx.float().sum(0)) < 1e-3, ("gelu_bwd fused bias-grad column sums", eng)
        if N % 64 == 0:
            dx = K.linear_dgrad(K.to_headmajor(dy), w, eng, dy_layout=L.LAYOUT_HEADMAJOR)
            assert K.rel_err(dx.float(), ref) < _tol(dtype), ("headmajor", eng)


@pytest.mark.parametrize("case", ENGINE_CASES, ids=lambda c: f"{c[0]}-{c[1]}")
@pytest.mark.parametrize("M,N,Kd", SHAPES)
def test_linear_wgrad(case, M, N, Kd):
    dtype, engines = _case(case)
    dy = randn(M, N, seed=31).to(dtype)
    x = randn(M, Kd, seed=32).to(dtype)
    ref_w = dy.float().t() @ x.float()
    ref_b = dy.float().sum(0)
    for eng in engines:
        dw, db = K.linear_wgrad(dy, x, N, Kd, eng)
        assert K.rel_err(dw, ref_w) < _tol(dtype), ("rowmajor", eng)
        assert K.rel_err(db, ref_b) < _tol(dtype)
        dw, db = K.linear_wgrad(K.to_headmajor(dy), x, N, Kd, eng, dy_layout=L.LAYOUT_HEADMAJOR)
        assert K.rel_err(dw, ref_w) < _tol(dtype), ("headmajor", eng)
        assert K.rel_err(db, ref_b) < _tol(dtype)


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("bn", [128, 192, 256])
def test_tcgen05_forced_block_n_and_streamk(bn, cg):
    """Every (BLOCK_N, CTA-group) instantiation of the tcgen05 kernel -- single CTAs (cta_group::1) and CTA pairs
    (cta_group::2, 256-row tiles) -- with and without stream-K, all epilogues that touch a second operand."""
    lib = L.load()
    M, N, Kd = 1576, 768, 768     # 6.16 tiles of 256 rows: partial last tile, second CTA of the last pair fully out of range
    x = randn(M, Kd, seed=91).to(torch.bfloat16)
    w = randn(N, Kd, seed=92, scale=0.05).to(torch.bfloat16)
    b = randn(N, seed=93, scale=0.5)
    dy = randn(M, N, seed=94).to(torch.bfloat16)
    res = randn(M, N, seed=95)
    u = randn(M, Kd, seed=96).to(torch.bfloat16)
    try:
        lib.vitk_debug_set(2, bn)
        lib.vitk_debug_set(4, cg)
        ref = x.float() @ w.float().t() + b
        y = K.linear_fwd(x, w, b, L.EPI_BIAS, L.ENGINE_TCGEN05)
        assert K.rel_err(y.float(), ref) < BF16_TOL
        y = K.linear_fwd(x, w, b, L.EPI_BIAS_RESIDUAL, L.ENGINE_TCGEN05, residual=res)
        assert K.rel_err(y, ref + res) < BF16_TOL
        y = K.linear_fwd(x, w, b, L.EPI_QKV_SCATTER, L.ENGINE_TCGEN05)
        assert K.rel_err(K.from_headmajor(y).float(), ref) < BF16_TOL
        dref = dy.float() @ w.float()
        dx = K.linear_dgrad(dy, w, L.ENGINE_TCGEN05)
        assert K.rel_err(dx.float(), dref) < BF16_TOL
        dx, cs = K.linear_dgrad(dy, w, L.ENGINE_TCGEN05, gelu_grad=u, want_colsum=True)
        assert K.rel_err(dx.float(), dref * u.float()) < BF16_TOL
        assert K.rel_err(cs, dx.float().sum(0)) < 1e-3
        dx = K.linear_dgrad(K.to_headmajor(dy), w, L.ENGINE_TCGEN05, dy_layout=L.LAYOUT_HEADMAJOR)
        assert K.rel_err(dx.float(), dref) < BF16_TOL
        # weight gradient: whole-K tiles (knob 1) / contiguous stream-K ranges (knob 13) / sliced split-K (default where
        # it fills the machine)
        for streamk_off, no_slices in ((0, 0), (0, 1), (1, 0)):
            lib.vitk_debug_set(1, streamk_off)
            lib.vitk_debug_set(13, no_slices)
            dw, _ = K.linear_wgrad(dy, x, N, Kd, L.ENGINE_TCGEN05)
            assert K.rel_err(dw, dy.float().t() @ x.float()) < BF16_TOL, (streamk_off, no_slices)
            dw, _ = K.linear_wgrad(K.to_headmajor(dy), x, N, Kd, L.ENGINE_TCGEN05, dy_layout=L.LAYOUT_HEADMAJOR)
            assert K.rel_err(dw, dy.float().t() @ x.float()) < BF16_TOL, (streamk_off, no_slices)
    finally:
        lib.vitk_debug_set(13, 0)
        lib.vitk_debug_set(1, 0)
        lib.vitk_debug_set(2, 0)
        lib.vitk_debug_set(4, 0)


# ------------------------------------------------------------------ attention
def _attn_ref(q, k, v):
    return F.scaled_dot_product_attention(q, k, v)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B", [1, 3, 26])   # 26: 312 (batch, head) items > 148 persistent CTAs -> prefetch path
def test_attention_fwd_bwd(dtype, B):
    M = B * 197
    qkv = randn(M, 2304, seed=41).to(dtype)
    dout = randn(M, 768, seed=42).to(dtype)
    hm = K.to_headmajor(qkv)
    out, lse = K.attn_fwd(hm, B)
    t = qkv.float().view(B, 197, 3, 12, 64).permute(2, 0, 3, 1, 4).clone().requires_grad_(True)
    o = _attn_ref(t[0], t[1], t[2]).transpose(1, 2).reshape(M, 768)
    assert K.rel_err(out.float(), o) < _tol(dtype)
    s = (t[0] @ t[1].transpose(-1, -2)) * 0.125
    lse_ref = torch.logsumexp(s, -1).permute(1, 0, 2).reshape(12, M)
    assert K.rel_err(lse, lse_ref) < _tol(dtype)
    o.backward(dout.float())
    dqkv, cs = K.attn_bwd(hm, out, dout, lse, B)
    ref = t.grad.permute(1, 3, 0, 2, 4).reshape(M, 2304)
    assert K.rel_err(K.from_headmajor(dqkv).float(), ref) < _tol(dtype) * (1 if dtype == torch.float32 else 2)
    assert K.rel_err(cs, K.from_headmajor(dqkv).float().sum(0)) < 1e-3     # fused qkv bias gradient


# ------------------------------------------------------------------ head
@pytest.mark.parametrize("B", [1, 5, 64])
@pytest.mark.parametrize("use_mask", [False, True])
def test_head_fwd_bwd(B, use_mask):
    C = 2
    feat = randn(B, 768, seed=51)
    ln_w, ln_b = 1 + 0.1 * randn(768, seed=52), 0.1 * randn(768, seed=53)
    w1, b1 = randn(512, 768, seed=54, scale=0.05), randn(512, seed=55, scale=0.1)
    w2, b2 = randn(C, 512, seed=56, scale=0.05), randn(C, seed=57, scale=0.1)
    m1 = m2 = None
    if use_mask:
        m1 = (torch.rand(B, 768, generator=_g(58), device=DEV) > 0.1).float() / 0.9
        m2 = (torch.rand(B, 512, generator=_g(59), device=DEV) > 0.1).float() / 0.9
    logits = torch.empty(B, C, device=DEV)
    save = torch.empty(L.load().vitk_head_save_floats(B), device=DEV)
    L.call("vitk_head_fwd", *[L.ptr(t) for t in (feat, ln_w, ln_b, w1, b1, w2, b2, m1, m2, logits, save)], B, C, L.stream_ptr())
    ps = [t.clone().requires_grad_(True) for t in (feat, ln_w, ln_b, w1, b1, w2, b2)]
    y = F.layer_norm(ps[0], (768,), ps[1], ps[2], 1e-5)
    if use_mask:
        y = y * m1
    h = F.gelu(F.linear(y, ps[3], ps[4]))
    if use_mask:
        h = h * m2
    ref = F.linear(h, ps[5], ps[6])
    assert K.rel_err(logits, ref) < FP32_TOL
    dl = randn(B, C, seed=60)
    ref.backward(dl)
    outs = [torch.zeros_like(t) for t in (feat, ln_w, ln_b, w1, b1, w2, b2)]
    L.call("vitk_head_bwd", L.ptr(dl), L.ptr(save), L.ptr(ln_w), L.ptr(w1), L.ptr(w2), L.ptr(m1), L.ptr(m2),
           *[L.ptr(t) for t in outs], B, C, L.stream_ptr())
    for got, p, name in zip(outs, ps, ["dfeat", "dln_w", "dln_b", "dw1", "db1", "dw2", "db2"]):
        assert K.rel_err(got, p.grad) < FP32_TOL * 2, name


# ------------------------------------------------------------------ focal loss vs the reference-generated golden
def test_focal_against_reference_golden(golden_dir):
    import vit_spoof_detection_pda_b200 as pkg
    g = torch.load(os.path.join(golden_dir, "focal_golden.pt"), weights_only=False)
    t = g["targets"].to(DEV)
    for case in g["cases"]:
        z = g["logits"].to(DEV).requires_grad_(True)
        crit = pkg.FocalLoss(case["alpha"], case["gamma"], case["reduction"])
        loss, met = crit(z, t, with_metrics=True)
        loss.backward()
        assert K.rel_err(loss.detach().cpu(), case["loss"]) < FP32_TOL, case
        assert K.rel_err(z.grad.cpu(), case["dlogits"]) < FP32_TOL, case
        assert torch.equal(met["preds"].cpu(), g["logits"].argmax(1))
        assert K.rel_err(met["probs_live"].cpu(), torch.softmax(g["logits"], 1)[:, 1]) < FP32_TOL
        assert int(met["ncorrect"].item()) == int((g["logits"].argmax(1) == g["targets"]).sum())
    none = pkg.FocalLoss(0.25, 2.0, "none")(g["logits"].to(DEV), t)
    assert K.rel_err(none.cpu(), g["none_reduction"]) < FP32_TOL


def test_focal_per_class_alpha():
    import vit_spoof_detection_pda_b200 as pkg
    from oracle import vit_oracle as vo
    z = randn(33, 2, seed=71, scale=2.0)
    t = torch.randint(0, 2, (33,), generator=_g(72), device=DEV)
    alpha = vo.class_weights_from_counts(n_live=8000, n_spoof=2500)
    zr = z.cpu().clone().requires_grad_(True)
    lr = vo.OracleFocalLoss(alpha, 2.0)(zr, t.cpu())
    lr.backward()
    zz = z.clone().requires_grad_(True)
    l = pkg.FocalLoss(alpha, 2.0)(zz, t)
    l.backward()
    assert K.rel_err(l.detach().cpu(), lr.detach()) < FP32_TOL
    assert K.rel_err(zz.grad.cpu(), zr.grad) < FP32_TOL


def test_focal_out_of_range_target_is_nan_not_oob():
    """A label outside [0, C) (F.cross_entropy device-asserts; ignore_index is not supported): that sample's loss and gradient
    become NaN and alpha[y] is never read out of bounds; the other samples are untouched (head_loss.cu)."""
    import vit_spoof_detection_pda_b200 as pkg
    z = randn(8, 2, seed=73, scale=2.0)
    t = torch.randint(0, 2, (8,), generator=_g(74), device=DEV)
    good = pkg.FocalLoss(0.25, 2.0, reduction="none")(z, t)
    for bad in (-100, 2, 7):
        tb = t.clone()
        tb[3] = bad
        zz = z.clone().requires_grad_(True)
        per = pkg.FocalLoss(0.25, 2.0, reduction="none")(zz, tb)
        per.sum().backward()
        assert torch.isnan(per[3]) and torch.isnan(zz.grad[3]).all()
        keep = [i for i in range(8) if i != 3]
        assert torch.equal(per[keep], good[keep]) and torch.isfinite(zz.grad[keep]).all()


# ------------------------------------------------------------------ fused Adam / AdamW / clip vs torch.optim
@pytest.mark.parametrize("adamw,lr,wd", [(True, 3e-4, 0.05), (False, 1e-5, 1e-4)])
@pytest.mark.parametrize("clip", [None, 1.0])
def test_adam_matches_torch(adamw, lr, wd, clip):
    n = 1_000_003
    p0 = randn(n, seed=81, scale=0.05)
    pr = p0.clone().requires_grad_(True)
    opt = (torch.optim.AdamW if adamw else torch.optim.Adam)([pr], lr=lr, weight_decay=wd, betas=(0.9, 0.999), eps=1e-8)
    p = p0.clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    p16 = torch.empty(n, dtype=torch.bfloat16, device=DEV)
    part = torch.empty(L.load().vitk_grad_sumsq_scratch_floats(), device=DEV)
    sumsq = torch.zeros(1, device=DEV)
    for step in range(1, 13):
        g = randn(n, seed=100 + step, scale=0.01 * step)
        pr.grad = g.clone()
        if clip is not None:
            torch.nn.utils.clip_grad_norm_([pr], clip)
            L.call("vitk_grad_sumsq", L.ptr(g), n, L.ptr(part), L.ptr(sumsq), L.stream_ptr())
            assert K.rel_err(sumsq.sqrt(), g.norm().reshape(1)) < 1e-5
        opt.step()
        L.call("vitk_adam_step", L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), L.ptr(p16), n, lr, 0.9, 0.999, 1e-8, wd,
               1 if adamw else 0, step, 1.0, L.ptr(sumsq) if clip is not None else None, clip or 0.0, L.stream_ptr())
    assert K.rel_err(p, pr.detach()) < 1e-5
    assert K.rel_err(m, opt.state[pr]["exp_avg"]) < 1e-5
    assert K.rel_err(v, opt.state[pr]["exp_avg_sq"]) < 1e-5
    assert torch.equal(p16, p.to(torch.bfloat16))
