"""One launch of every libvitk kernel family at tiny shapes -- the command compute-sanitizer wraps:

    compute-sanitizer --tool memcheck  --kernel-name-exclude regex:at:: python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck ...        compute-sanitizer --tool synccheck ...

Covers: every (BLOCK_N, CTA-group) instantiation of the tcgen05 GEMM with each epilogue kind (store, GELU, residual,
head-major scatter, GELU' multiply + column sums, accumulate: whole-K / stream-K / sliced split-K), both operand
major-nesses and the head-major operand maps; attention forward / backward (tcgen05 and FFMA); LayerNorm forward /
backward; classifier head; focal loss; sum of squares / gradient scale / Adam / bf16 cast; patch embedding forward and
weight gradient; the SIMT GEMM.  Results are compared with torch so a sanitizer-clean run is also a correct one."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import kernels_api as K  # noqa: E402
import vit_spoof_detection_pda_b200 as pkg  # noqa: E402
from vit_spoof_detection_pda_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
lib = L.load()
g = torch.Generator(device="cuda").manual_seed(0)


def rn(*s, scale=1.0, dt=torch.float32):
    return (torch.randn(*s, generator=g, device=dev) * scale).to(dt)


def ok(name, err, tol):
    print(f"{name:60s} err {err:.3e}", flush=True)
    assert err < tol, name


E = L.ENGINE_TCGEN05
bf = torch.bfloat16
M, N, Kd = 394, 768, 768
x, w, b = rn(M, Kd, dt=bf), rn(N, Kd, scale=0.05, dt=bf), rn(N, scale=0.5)
dy, res, u = rn(M, N, dt=bf), rn(M, N), rn(M, Kd, dt=bf)
ref = x.float() @ w.float().t() + b
dref = dy.float() @ w.float()
wref = dy.float().t() @ x.float()
for cg in (1, 2):
    for bn in (128, 192, 256):
        lib.vitk_debug_set(2, bn)
        lib.vitk_debug_set(4, cg)
        t = f"gemm_tc BN{bn} CG{cg} "
        ok(t + "store", K.rel_err(K.linear_fwd(x, w, b, L.EPI_BIAS, E).float(), ref), 2e-2)
        ok(t + "residual", K.rel_err(K.linear_fwd(x, w, b, L.EPI_BIAS_RESIDUAL, E, residual=res), ref + res), 2e-2)
        ok(t + "scatter", K.rel_err(K.from_headmajor(K.linear_fwd(x, w, b, L.EPI_QKV_SCATTER, E)).float(), ref), 2e-2)
        gl, dgl = K.linear_fwd(x, w, b, L.EPI_BIAS_GELU, E)
        ok(t + "gelu", K.rel_err(gl.float(), F.gelu(ref.to(bf).float())), 2e-2)
        ok(t + "dgrad", K.rel_err(K.linear_dgrad(dy, w, E).float(), dref), 2e-2)
        dx, cs = K.linear_dgrad(dy, w, E, gelu_grad=u, want_colsum=True)
        ok(t + "dgrad x gelu' + colsum", max(K.rel_err(dx.float(), dref * u.float()), K.rel_err(cs, dx.float().sum(0))), 2e-2)
        ok(t + "dgrad head-major", K.rel_err(K.linear_dgrad(K.to_headmajor(dy), w, E, dy_layout=L.LAYOUT_HEADMAJOR).float(), dref), 2e-2)
        for k1, k13 in ((0, 0), (0, 1), (1, 0)):
            lib.vitk_debug_set(1, k1)
            lib.vitk_debug_set(13, k13)
            ok(t + f"wgrad mode({k1},{k13})", K.rel_err(K.linear_wgrad(dy, x, N, Kd, E)[0], wref), 2e-2)
            ok(t + f"wgrad head-major mode({k1},{k13})",
               K.rel_err(K.linear_wgrad(K.to_headmajor(dy), x, N, Kd, E, dy_layout=L.LAYOUT_HEADMAJOR)[0], wref), 2e-2)
        lib.vitk_debug_set(1, 0)
        lib.vitk_debug_set(13, 0)
lib.vitk_debug_set(2, 0)
lib.vitk_debug_set(4, 0)
x32, w32, dy32 = x.float()[:200], w.float(), dy.float()[:200]
ok("gemm_simt fwd", K.rel_err(K.linear_fwd(x32, w32, b, L.EPI_BIAS, L.ENGINE_SIMT), x32 @ w32.t() + b), 1e-4)
ok("gemm_simt wgrad", K.rel_err(K.linear_wgrad(dy32, x32, N, Kd, L.ENGINE_SIMT)[0], dy32.t() @ x32), 1e-4)

for B in (1, 3):
    Mb = B * 197
    for dt, tol in ((bf, 3e-2), (torch.float32, 1e-4)):
        qkv = rn(Mb, 2304, dt=dt)
        dout = rn(Mb, 768, dt=dt)
        q, k, v = [t.reshape(B, 197, 12, 64).permute(0, 2, 1, 3).float().requires_grad_(True) for t in qkv.split(768, dim=1)]
        o = F.scaled_dot_product_attention(q, k, v)
        o.backward(dout.float().reshape(B, 197, 12, 64).permute(0, 2, 1, 3))
        out, lse = K.attn_fwd(K.to_headmajor(qkv), B)
        ok(f"attention fwd B{B} {dt}", K.rel_err(out.float(), o.permute(0, 2, 1, 3).reshape(Mb, 768)), tol)
        dqkv, cs = K.attn_bwd(K.to_headmajor(qkv), out, dout, lse, B)
        rd = torch.cat([t.grad.permute(0, 2, 1, 3).reshape(Mb, 768) for t in (q, k, v)], dim=1)
        ok(f"attention bwd B{B} {dt}", K.rel_err(K.from_headmajor(dqkv).float(), rd), tol)

xx = rn(300, 768, scale=2.0) + 0.5
gam, bet = 1 + 0.1 * rn(768), 0.1 * rn(768)
for odt in (torch.float32, bf):
    y, mean, rstd = K.layernorm_fwd(xx, gam, bet, 1e-6, odt)
    ok(f"layernorm fwd {odt}", K.rel_err(y.float(), F.layer_norm(xx, (768,), gam, bet, 1e-6)), 2e-2)
    dyl = rn(300, 768).to(odt)
    dxl, dx16, dg, db, cs = K.layernorm_bwd(dyl, xx, gam, mean, rstd, dres=rn(300, 768), want16=True)
    ok(f"layernorm bwd {odt} (finite)", float(~torch.isfinite(dxl).all()), 0.5)

# whole model (head, focal, patch embedding, embed grads, scatter, adam, sumsq, cast) at depth 1, both precisions
for precision in ("fp32", "bf16"):
    m = pkg.ViTFaceAntiSpoofing(dropout=0.1, depth=1, precision=precision).to(dev).train()
    opt = pkg.FusedAdam(m.parameters(), lr=1e-4, weight_decay=1e-4, adamw=False)
    img = rn(2, 3, 224, 224)
    lab = torch.tensor([0, 1], device=dev)
    for _ in range(2):
        loss, met = pkg.FocalLoss(0.25, 2.0)(m(img), lab, with_metrics=True)
        loss.backward()
        pkg.clip_grad_norm_(m.parameters(), 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
    m.eval()
    with torch.no_grad():
        lo = m(img)
        lo8 = m((torch.rand(2, 224, 224, 3, generator=g, device=dev) * 255).to(torch.uint8))
    ok(f"model step + eval {precision} (finite)", float(~(torch.isfinite(lo).all() & torch.isfinite(lo8).all() & torch.isfinite(loss))), 0.5)
torch.cuda.synchronize()
print("sanitize_smoke ok")
