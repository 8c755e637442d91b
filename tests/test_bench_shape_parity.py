"""Parity at the BENCHMARKED configuration (BASELINE configs[1]: depth 12, batch 64 -> M = 12,608 token rows), where the
sliced split-K weight gradients, the 2-wave forward / dgrad tile grids and the full-size attention grids take the
decompositions bench.py times ("missing 2"):

  * bf16 path: logits and ALL 156 gradients against the fp32 validation path of the same weights
    (north_star: <= 2e-2 max-abs; and <= 0.1 of each tensor's scale);
  * fp32 validation path against the oracle executed on the GPU in fp32 with TF32 off (<= 1e-4 relative);
  * the twelve GEMM shapes of one training step (profiles/r1_cublas_yardstick.md) through the C-ABI against
    x.float() @ w.float().T, including the head-major qkv operands and every fused epilogue.
Reference call sites: /root/reference/train_advanced.py:322-338 (step), :203 (encoder forward).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import vit_oracle as vo  # checker only

if torch.cuda.is_available():
    import vit_spoof_detection_pda_b200 as pkg
    from vit_spoof_detection_pda_b200 import _lib as L
    import kernels_api as K
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    DEV = torch.device("cuda:0")
else:
    pkg = L = K = DEV = None

B = 64
M = B * 197


def _pair(precision, seed=42):
    ref = vo.OracleViTFaceAntiSpoofing(dropout=0.0, depth=12)
    vo.seeded_init_(ref, seed=seed)
    m = pkg.ViTFaceAntiSpoofing(dropout=0.0, depth=12, precision=precision)
    m.load_state_dict(ref.state_dict(), strict=True)
    return ref, m.to(DEV).train()


def _fwd_bwd(m, images, labels):
    out = m(images)
    loss = pkg.FocalLoss(0.25, 2.0)(out, labels)
    loss.backward()
    torch.cuda.synchronize()
    return out.detach().clone(), {n: p.grad.detach().clone() for n, p in m.named_parameters()}


def test_bf16_step_bs64_all_gradients_vs_fp32_validation_path():
    images, labels = vo.synthetic_batch(B, seed=21)
    images, labels = images.to(DEV), labels.to(DEV)
    _, m32 = _pair("fp32")
    out32, g32 = _fwd_bwd(m32, images, labels)
    del m32
    torch.cuda.empty_cache()
    _, m16 = _pair("bf16")
    out16, g16 = _fwd_bwd(m16, images, labels)
    err = float((out16 - out32).abs().max())
    assert err < 2e-2, err
    margin = (out32[:, 1] - out32[:, 0]).abs()
    decided = margin > 2 * err
    assert torch.equal(out16.argmax(1)[decided], out32.argmax(1)[decided])
    assert len(g16) == 156
    worst_abs, worst_rel = ("", 0.0), ("", 0.0)
    for n in g32:
        a, b = g16[n].double(), g32[n].double()
        e_abs = float((a - b).abs().max())
        e_rel = e_abs / max(float(b.abs().max()), 1e-30)
        if e_abs > worst_abs[1]:
            worst_abs = (n, e_abs)
        if e_rel > worst_rel[1]:
            worst_rel = (n, e_rel)
    print(f"bs64 bf16 vs fp32-validate: logits max-abs {err:.3e}; grads worst max-abs {worst_abs}, worst rel-to-scale {worst_rel}")
    assert worst_abs[1] < 2e-2, worst_abs
    assert worst_rel[1] < 0.1, worst_rel


def test_fp32_step_bs64_vs_oracle_on_gpu():
    ref, m = _pair("fp32")
    ref = ref.to(DEV).train()
    images, labels = vo.synthetic_batch(B, seed=22)
    images, labels = images.to(DEV), labels.to(DEV)
    out_r = ref(images)
    vo.OracleFocalLoss(0.25, 2.0)(out_r, labels).backward()
    out, g = _fwd_bwd(m, images, labels)
    scale = float(out_r.abs().max())
    assert float((out - out_r).abs().max()) <= 1e-4 * scale
    worst = ("", 0.0)
    for n, q in ref.named_parameters():
        e = float((g[n].double() - q.grad.double()).abs().max() / q.grad.double().abs().max().clamp_min(1e-30))
        if e > worst[1]:
            worst = (n, e)
    print(f"bs64 fp32-validate vs oracle(GPU, TF32 off): worst gradient rel err {worst}")
    assert worst[1] < 1e-4, worst


def _randn(*shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, generator=g, device=DEV) * scale


# (name, N = out features, K = in features) of the four Linear layers of a block
LINEARS = [("qkv", 2304, 768), ("proj", 768, 768), ("fc1", 3072, 768), ("fc2", 768, 3072)]


@pytest.mark.parametrize("name,N,Kd", LINEARS, ids=[l[0] for l in LINEARS])
def test_bench_gemm_shapes_m12608(name, N, Kd):
    """forward (with the epilogue the model uses), dgrad and wgrad of each Linear at M = 12,608 on the tcgen05 engine."""
    E = L.ENGINE_TCGEN05
    x = _randn(M, Kd, seed=31).to(torch.bfloat16)
    w = _randn(N, Kd, seed=32, scale=0.03).to(torch.bfloat16)
    b = _randn(N, seed=33, scale=0.5)
    dy = _randn(M, N, seed=34).to(torch.bfloat16)
    ref = x.float() @ w.float().t() + b
    tol = 2e-2
    if name == "qkv":
        y = K.linear_fwd(x, w, b, L.EPI_QKV_SCATTER, E)
        assert K.rel_err(K.from_headmajor(y).float(), ref) < tol
    elif name == "fc1":
        g, dg = K.linear_fwd(x, w, b, L.EPI_BIAS_GELU, E)
        ur = ref.to(torch.bfloat16).float().requires_grad_(True)
        gr = F.gelu(ur)
        gr.sum().backward()
        assert K.rel_err(g.float(), gr) < tol and K.rel_err(dg.float(), ur.grad) < tol
    else:
        res = _randn(M, N, seed=35)
        y = K.linear_fwd(x, w, b, L.EPI_BIAS_RESIDUAL, E, residual=res)
        assert K.rel_err(y, ref + res) < tol
    del ref
    # dgrad: dX = dY W  (qkv: head-major dY; fc2: x GELU' with the fused fc1 bias-gradient column sums; proj: fused column sums)
    dref = dy.float() @ w.float()
    if name == "qkv":
        dx = K.linear_dgrad(K.to_headmajor(dy), w, E, dy_layout=L.LAYOUT_HEADMAJOR)
        assert K.rel_err(dx.float(), dref) < tol
    elif name == "fc2":
        u = _randn(M, Kd, seed=36).to(torch.bfloat16)
        dx, cs = K.linear_dgrad(dy, w, E, gelu_grad=u, want_colsum=True)
        assert K.rel_err(dx.float(), dref * u.float()) < tol
        assert K.rel_err(cs, dx.float().sum(0)) < 2e-3
    elif name == "proj":
        dx, cs = K.linear_dgrad(dy, w, E, want_colsum=True)
        assert K.rel_err(dx.float(), dref) < tol
        assert K.rel_err(cs, dx.float().sum(0)) < 2e-3
    else:
        dx = K.linear_dgrad(dy, w, E)
        assert K.rel_err(dx.float(), dref) < tol
    del dref
    # wgrad: dW = dY^T X over 12,608 rows (sliced split-K where it fills the machine, stream-K ranges for qkv)
    wref = dy.float().t() @ x.float()
    if name == "qkv":
        dw, _ = K.linear_wgrad(K.to_headmajor(dy), x, N, Kd, E, dy_layout=L.LAYOUT_HEADMAJOR)
    else:
        dw, _ = K.linear_wgrad(dy, x, N, Kd, E)
    assert K.rel_err(dw, wref) < 5e-3      # fp32 accumulation of exact bf16 products: only the summation order differs


def _tail_plan(I, J, R, b_mn, f32_out=0):
    import ctypes as C
    nw, nt, r0 = C.c_int(-1), C.c_int(-1), C.c_int(-1)
    assert L.load().vitk_gemm_tail_plan(I, J, R, b_mn, f32_out, C.byref(nw), C.byref(nt), C.byref(r0)) == 0
    return nw.value, nt.value, r0.value


def _tail_scratch():
    return torch.zeros(L.load().vitk_gemm_tail_scratch_floats(768), dtype=torch.float32, device=DEV)


@pytest.mark.parametrize("which", ["fc2_fwd", "fc1_dgrad", "qkv_dgrad", "bf16_fwd"])
def test_split_tail_m12608(which):
    """The deep J = 768 GEMMs of a bs-64 step with the split tail (vitk_linear_*_ws: 148 whole tiles with the fused
    epilogue, the k-blocks of the last two tiles dealt out to all 74 CTA pairs, last-arriver epilogue) against the fp32 product
    AND against the plain launch; the scratch (tickets + partial sums) comes back all zero, five times in a row -- it is
    reused by the next GEMM of the step."""
    E = L.ENGINE_TCGEN05
    scratch = _tail_scratch()
    if which == "fc2_fwd":
        x = _randn(M, 3072, seed=51).to(torch.bfloat16)
        w = _randn(768, 3072, seed=52, scale=0.03).to(torch.bfloat16)
        b = _randn(768, seed=53, scale=0.5)
        res = _randn(M, 768, seed=54)
        plan = _tail_plan(M, 768, 3072, 0, 1)
        ref = x.float() @ w.float().t() + b + res
        run = lambda s: K.linear_fwd(x, w, b, L.EPI_BIAS_RESIDUAL, E, residual=res, scratch=s)   # noqa: E731
    elif which == "bf16_fwd":
        x = _randn(M, 3072, seed=51).to(torch.bfloat16)
        w = _randn(768, 3072, seed=52, scale=0.03).to(torch.bfloat16)
        b = _randn(768, seed=53, scale=0.5)
        plan = _tail_plan(M, 768, 3072, 0)
        ref = x.float() @ w.float().t() + b
        run = lambda s: K.linear_fwd(x, w, b, L.EPI_BIAS, E, scratch=s)   # noqa: E731
    elif which == "fc1_dgrad":
        dy = _randn(M, 3072, seed=55).to(torch.bfloat16)
        w = _randn(3072, 768, seed=56, scale=0.03).to(torch.bfloat16)
        plan = _tail_plan(M, 768, 3072, 1)
        ref = dy.float() @ w.float()
        run = lambda s: K.linear_dgrad(dy, w, E, scratch=s)   # noqa: E731
    else:
        dy = _randn(M, 2304, seed=57).to(torch.bfloat16)
        w = _randn(2304, 768, seed=58, scale=0.03).to(torch.bfloat16)
        dyh = K.to_headmajor(dy)
        plan = _tail_plan(M, 768, 2304, 1)
        ref = dy.float() @ w.float()
        run = lambda s: K.linear_dgrad(dyh, w, E, dy_layout=L.LAYOUT_HEADMAJOR, scratch=s)   # noqa: E731
    assert plan == (148, 2, 49 * 256)
    r0 = plan[2]
    single = run(None).float()
    for _ in range(5):
        split = run(scratch).float()
        torch.cuda.synchronize()
        assert int(torch.count_nonzero(scratch)) == 0
        assert K.rel_err(split, ref) < 2e-2
        assert K.rel_err(split[r0:], ref[r0:]) < 2e-2
        # same arithmetic up to the fp32 summation order of the tail tiles' k-ranges (and one bf16 rounding step where that flips)
        assert float((split - single).abs().max()) <= 2e-2 * float(ref.abs().max())
        assert float((split[r0:] - single[r0:]).abs().mean()) <= 2e-3 * float(ref.abs().mean())
        print(f"{which}: whole-tile rows bit-identical to the plain launch: {bool(torch.equal(split[:r0], single[:r0]))}")


@pytest.mark.parametrize("batch", [32, 33, 49, 50])
def test_split_tail_fp32_residual_epilogue(batch):
    """fc2 forward (bias + fp32 residual, fp32 out) at other tail geometries: 256 x 256 tiles (single fp32 staging slot) with
    1 / 4 tail tiles (bs 32 / 33), 256 x 192 tiles (two slots) with 4 / 8 tail tiles (bs 49 / 50) -- cluster ranges that
    straddle tail tiles, tails that start in the middle of a tile row."""
    E = L.ENGINE_TCGEN05
    Mb = batch * 197
    nw, nt, r0 = _tail_plan(Mb, 768, 3072, 0, 1)
    assert nt > 0
    scratch = _tail_scratch()
    x = _randn(Mb, 3072, seed=61).to(torch.bfloat16)
    w = _randn(768, 3072, seed=62, scale=0.03).to(torch.bfloat16)
    b = _randn(768, seed=63, scale=0.5)
    res = _randn(Mb, 768, seed=64)
    ref = x.float() @ w.float().t() + b + res
    single = K.linear_fwd(x, w, b, L.EPI_BIAS_RESIDUAL, E, residual=res)
    for _ in range(3):
        y = K.linear_fwd(x, w, b, L.EPI_BIAS_RESIDUAL, E, residual=res, scratch=scratch)
        torch.cuda.synchronize()
        assert int(torch.count_nonzero(scratch)) == 0
        assert K.rel_err(y, ref) < 2e-2
        assert K.rel_err(y, single) < 1e-4       # fp32 outputs: only the summation order of the tail tiles differs


@pytest.mark.parametrize("batch", [16, 17, 33, 66])
def test_split_tail_dgrad_other_batches(batch):
    """fc1 dgrad (row-major dY) and qkv dgrad (head-major dY) with other tail geometries: 256 x 128 tiles with 4 / 10 tail tiles
    (bs 16 / 17), 256 x 256 tiles with 4 / 5 tail tiles (bs 33 / 66)."""
    E = L.ENGINE_TCGEN05
    Mb = batch * 197
    scratch = _tail_scratch()
    assert _tail_plan(Mb, 768, 3072, 1)[1] > 0
    dy = _randn(Mb, 3072, seed=65).to(torch.bfloat16)
    w = _randn(3072, 768, seed=66, scale=0.03).to(torch.bfloat16)
    dref = dy.float() @ w.float()
    single = K.linear_dgrad(dy, w, E).float()
    for _ in range(3):
        dx = K.linear_dgrad(dy, w, E, scratch=scratch).float()
        assert K.rel_err(dx, dref) < 2e-2
        assert float((dx - single).abs().max()) <= 2e-2 * float(dref.abs().max())
    if _tail_plan(Mb, 768, 2304, 1)[1] > 0:
        dy = _randn(Mb, 2304, seed=67).to(torch.bfloat16)
        wq = _randn(2304, 768, seed=68, scale=0.03).to(torch.bfloat16)
        dyh = K.to_headmajor(dy)
        dref = dy.float() @ wq.float()
        for _ in range(3):
            dx = K.linear_dgrad(dyh, wq, E, dy_layout=L.LAYOUT_HEADMAJOR, scratch=scratch).float()
            assert K.rel_err(dx, dref) < 2e-2
    torch.cuda.synchronize()
    assert int(torch.count_nonzero(scratch)) == 0


def test_split_tail_shallow_reductions_behind_knob():
    """vitk_debug_set(9, 12) lowers the depth threshold of the split tail from 24 to 12 k-blocks (A/B knob): qkv forward (head-
    major scatter epilogue, 450 tiles = 6 waves + 6 tiles) and fc1 forward (bias + GELU, two outputs, 600 tiles = 8 waves + 8
    tiles) then take it too -- the last arriver's per-thread epilogue implements every mode."""
    E = L.ENGINE_TCGEN05
    lib = L.load()
    scratch = _tail_scratch()
    x = _randn(M, 768, seed=71).to(torch.bfloat16)
    lib.vitk_debug_set(9, 12)
    try:
        assert _tail_plan(M, 2304, 768, 0)[1] == 6 and _tail_plan(M, 3072, 768, 0)[1] == 8
        w = _randn(2304, 768, seed=72, scale=0.03).to(torch.bfloat16)
        b = _randn(2304, seed=73, scale=0.5)
        ref = x.float() @ w.float().t() + b
        plain = K.from_headmajor(K.linear_fwd(x, w, b, L.EPI_QKV_SCATTER, E)).float()
        for _ in range(3):
            y = K.from_headmajor(K.linear_fwd(x, w, b, L.EPI_QKV_SCATTER, E, scratch=scratch)).float()
            assert K.rel_err(y, ref) < 2e-2 and float((y - plain).abs().max()) <= 2e-2 * float(ref.abs().max())
        w1 = _randn(3072, 768, seed=74, scale=0.03).to(torch.bfloat16)
        b1 = _randn(3072, seed=75, scale=0.5)
        g0, d0 = K.linear_fwd(x, w1, b1, L.EPI_BIAS_GELU, E)
        for _ in range(3):
            g, dg = K.linear_fwd(x, w1, b1, L.EPI_BIAS_GELU, E, scratch=scratch)
            # the tail rows apply erf-GELU to the same bf16-rounded pre-activation the fused epilogue's fast GELU sees
            assert float((g.float() - g0.float()).abs().max()) <= 2e-2 * float(g0.float().abs().max())
            assert float((dg.float() - d0.float()).abs().max()) <= 2e-2
        torch.cuda.synchronize()
        assert int(torch.count_nonzero(scratch)) == 0
    finally:
        lib.vitk_debug_set(9, 0)


def test_attention_bs64_vs_sdpa():
    """the full-size attention grids (768 (batch, head) items on 148 persistent CTAs: 5.2 items per CTA)."""
    qkv = _randn(M, 2304, seed=41, scale=1.0).to(torch.bfloat16)
    dout = _randn(M, 768, seed=42).to(torch.bfloat16)
    q, k, v = [t.reshape(B, 197, 12, 64).permute(0, 2, 1, 3).float().requires_grad_(True) for t in qkv.split(768, dim=1)]
    o = F.scaled_dot_product_attention(q, k, v)
    o.backward(dout.float().reshape(B, 197, 12, 64).permute(0, 2, 1, 3))
    out, lse = K.attn_fwd(K.to_headmajor(qkv), B)
    ref_o = o.permute(0, 2, 1, 3).reshape(M, 768)
    assert K.rel_err(out.float(), ref_o) < 2e-2
    dqkv, cs = K.attn_bwd(K.to_headmajor(qkv), out, dout, lse, B)
    ref_d = torch.cat([t.grad.permute(0, 2, 1, 3).reshape(M, 768) for t in (q, k, v)], dim=1)
    assert K.rel_err(K.from_headmajor(dqkv).float(), ref_d) < 3e-2
    assert K.rel_err(cs, K.from_headmajor(dqkv).float().sum(0)) < 5e-3
