"""Direct parity of the im2col-free patch embedding (csrc/patch_embed.cu) through the C-ABI against F.conv2d in fp32
(timm PatchEmbed.proj = Conv2d(3, 768, k16, s16), reached from /root/reference/train_advanced.py:190/203; SURVEY.md 8a
a2.1/a2.2, 8f n2): forward from fp32 NCHW, forward from uint8 HWC (ToTensor + Normalize inside the loader), and the
weight gradient, on the tcgen05 path (TMA boxes of the image -> converter warps -> tcgen05.mma) and on the fp32-validate path.
Tolerances: bf16 operands with fp32 accumulation over K = 768 -> 2e-2 of the output scale (north_star); fp32: 1e-4."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from vit_spoof_detection_pda_b200 import _lib as L
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    DEV = torch.device("cuda:0")
else:
    L = DEV = None

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def _rn(*shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, generator=g, device=DEV) * scale


def _params(seed=0):
    w = _rn(768, 3, 16, 16, seed=seed + 1, scale=0.03)
    b = _rn(768, seed=seed + 2, scale=0.3)
    cls = _rn(768, seed=seed + 3, scale=0.5)
    pos = _rn(197, 768, seed=seed + 4, scale=0.5)
    return w, b, cls, pos


def _ref_fwd(img, w, b, cls, pos):
    y = F.conv2d(img, w, b, stride=16)                         # [B, 768, 14, 14]
    tok = y.flatten(2).transpose(1, 2)                         # [B, 196, 768]
    x0 = torch.cat([cls.expand(img.shape[0], 1, 768), tok], dim=1) + pos
    return x0.reshape(-1, 768)


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


CASES = [("bf16", 2e-2), ("fp32", 1e-4)]


@pytest.mark.parametrize("batch", [1, 3, 64])
@pytest.mark.parametrize("prec,tol", CASES, ids=[c[0] for c in CASES])
def test_patch_embed_fwd_vs_conv2d(batch, prec, tol):
    w, b, cls, pos = _params()
    img = _rn(batch, 3, 224, 224, seed=10 + batch)
    x0 = torch.full((batch * 197, 768), float("nan"), device=DEV)
    precision = L.PREC_BF16 if prec == "bf16" else L.PREC_FP32
    w16 = w.to(torch.bfloat16)
    L.call("vitk_patch_embed_fwd", L.ptr(img), L.ptr(w), L.ptr(w16), L.ptr(b), L.ptr(cls), L.ptr(pos), L.ptr(x0), batch, precision,
           L.ENGINE_AUTO, L.stream_ptr())
    ref = _ref_fwd(img, w, b, cls, pos)
    assert torch.isfinite(x0).all()
    assert _rel(x0, ref) < tol
    # the CLS rows are exact in every precision (no product involved)
    assert torch.equal(x0.view(batch, 197, 768)[:, 0], (cls + pos[0]).expand(batch, 768))


@pytest.mark.parametrize("batch", [1, 5])
@pytest.mark.parametrize("prec,tol", CASES, ids=[c[0] for c in CASES])
def test_patch_embed_fwd_u8_vs_normalised_fp32(batch, prec, tol):
    """uint8 HWC pixels through the converter stage == ToTensor + Normalize in torch, then conv2d."""
    w, b, cls, pos = _params(seed=20)
    g = torch.Generator(device="cuda").manual_seed(5)
    u8 = torch.randint(0, 256, (batch, 224, 224, 3), generator=g, device=DEV, dtype=torch.uint8)
    # ToTensor + Normalize as the reference's transforms run them: on the CPU (true division; torch's CUDA kernel would
    # multiply by the reciprocal of 255)
    mean = torch.tensor(MEAN).view(1, 3, 1, 1)
    std = torch.tensor(STD).view(1, 3, 1, 1)
    img = ((u8.cpu().permute(0, 3, 1, 2).to(torch.float32).div(255.0) - mean) / std).contiguous().to(DEV)
    import ctypes as C
    m3, s3 = (C.c_float * 3)(*MEAN), (C.c_float * 3)(*STD)
    x0 = torch.full((batch * 197, 768), float("nan"), device=DEV)
    scratch = torch.empty(batch, 3, 224, 224, device=DEV)
    precision = L.PREC_BF16 if prec == "bf16" else L.PREC_FP32
    w16 = w.to(torch.bfloat16)
    L.call("vitk_patch_embed_fwd_u8", L.ptr(u8), m3, s3, L.ptr(w), L.ptr(w16), L.ptr(b), L.ptr(cls), L.ptr(pos), L.ptr(x0),
           L.ptr(scratch), batch, precision, L.ENGINE_AUTO, L.stream_ptr())
    assert _rel(x0, _ref_fwd(img, w, b, cls, pos)) < tol
    # and it is THE SAME arithmetic as the fp32-NCHW entry on the normalised image (identical operand values, same kernel)
    x1 = torch.empty_like(x0)
    imgc = img.contiguous()
    L.call("vitk_patch_embed_fwd", L.ptr(imgc), L.ptr(w), L.ptr(w16), L.ptr(b), L.ptr(cls), L.ptr(pos), L.ptr(x1), batch, precision,
           L.ENGINE_AUTO, L.stream_ptr())
    assert torch.equal(x0, x1)
    if prec == "fp32":   # the stand-alone ToTensor + Normalize entry
        out = torch.empty(batch, 3, 224, 224, device=DEV)
        L.call("vitk_u8_to_nchw", L.ptr(u8), m3, s3, L.ptr(out), batch, L.stream_ptr())
        assert torch.equal(out, imgc)


@pytest.mark.parametrize("batch", [1, 3, 64])
@pytest.mark.parametrize("prec,tol", CASES, ids=[c[0] for c in CASES])
def test_patch_embed_wgrad_vs_conv2d_autograd(batch, prec, tol):
    w, b, cls, pos = _params(seed=30)
    img = _rn(batch, 3, 224, 224, seed=40 + batch)
    dx0 = _rn(batch * 197, 768, seed=50 + batch)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    cr, pr = cls.clone().requires_grad_(True), pos.clone().requires_grad_(True)
    (_ref_fwd(img, wr, br, cr, pr) * dx0).sum().backward()
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    dcls, dpos = torch.zeros_like(cls), torch.zeros_like(pos)
    precision = L.PREC_BF16 if prec == "bf16" else L.PREC_FP32
    dx16 = dx0.to(torch.bfloat16)
    L.call("vitk_patch_embed_wgrad", L.ptr(dx0), L.ptr(dx16), L.ptr(img), L.ptr(dw), L.ptr(db), L.ptr(dcls), L.ptr(dpos), batch,
           precision, L.ENGINE_AUTO, L.stream_ptr())
    assert _rel(dw, wr.grad) < tol
    assert _rel(db, br.grad) < 1e-4 and _rel(dcls, cr.grad) < 1e-4 and _rel(dpos, pr.grad) < 1e-4
    # gradients ACCUMULATE (+=): a second call doubles them
    L.call("vitk_patch_embed_wgrad", L.ptr(dx0), L.ptr(dx16), L.ptr(img), L.ptr(dw), L.ptr(db), L.ptr(dcls), L.ptr(dpos), batch,
           precision, L.ENGINE_AUTO, L.stream_ptr())
    assert _rel(dw, 2 * wr.grad) < tol and _rel(dpos, 2 * pr.grad) < 1e-4
