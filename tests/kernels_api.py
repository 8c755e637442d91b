"""Thin torch-tensor wrappers over the C-ABI used by the GPU parity tests (calls go through ctypes,
exactly as the product's Python host code does)."""
import torch

from vit_spoof_detection_pda_b200 import _lib as L

DT = {torch.float32: L.F32, torch.bfloat16: L.BF16}


class GuardArena:
    """Output buffers carved from one sentinel-filled device arena with guard bands on both sides: after the kernels ran,
    check() proves that nothing was written outside the buffers they were given (compute-sanitizer's memcheck is closed on
    this GPU pool; tests/test_guard_bands.py).  Installed with `kernels_api.GUARD = GuardArena(...)`."""
    BYTE = 0xA5

    def __init__(self, device, nbytes=512 << 20, guard=64 << 10):
        self.buf = torch.full((nbytes,), self.BYTE, dtype=torch.uint8, device=device)
        self.guard, self.cur, self.spans = guard, guard, []

    def alloc(self, shape, dtype, zero):
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        off = (self.cur + 255) // 256 * 256
        assert off + nbytes + self.guard <= self.buf.numel(), "guard arena too small"
        self.spans.append((off, nbytes))
        self.cur = off + nbytes + self.guard
        t = self.buf[off:off + nbytes].view(dtype).view(*shape)
        if zero:
            t.zero_()
        return t

    def check(self):
        """every byte that is not payload still holds the sentinel; returns the number of bytes checked"""
        if self.buf.is_cuda:
            torch.cuda.synchronize()
        checked, prev = 0, 0
        for off, nbytes in self.spans + [(self.buf.numel(), 0)]:
            gap = self.buf[prev:off]
            bad = int((gap != self.BYTE).sum())
            assert bad == 0, f"{bad} bytes written outside the output buffers (arena offsets {prev}..{off})"
            checked += gap.numel()
            prev = off + nbytes
        return checked


GUARD = None


def _empty(*shape, dtype, device):
    return torch.empty(*shape, dtype=dtype, device=device) if GUARD is None else GUARD.alloc(shape, dtype, False)


def _zeros(*shape, dtype, device):
    return torch.zeros(*shape, dtype=dtype, device=device) if GUARD is None else GUARD.alloc(shape, dtype, True)


def to_headmajor(x):  # [M, C] -> [C/64, M, 64]
    M, Cc = x.shape
    return x.reshape(M, Cc // 64, 64).permute(1, 0, 2).contiguous()


def from_headmajor(x):  # [C/64, M, 64] -> [M, C]
    nb, M, _ = x.shape
    return x.permute(1, 0, 2).reshape(M, nb * 64).contiguous()


def layernorm_fwd(x, gamma, beta, eps, out_dtype, x_stride=None, rows=None):
    rows = x.shape[0] if rows is None else rows
    y = _empty(rows, 768, dtype=out_dtype, device=x.device)
    mean = _empty(rows, dtype=torch.float32, device=x.device)
    rstd = _empty(*mean.shape, dtype=mean.dtype, device=mean.device)
    L.call("vitk_layernorm_fwd", L.ptr(x), x_stride or 768, L.ptr(gamma), L.ptr(beta), L.ptr(y), DT[out_dtype],
           L.ptr(mean), L.ptr(rstd), rows, eps, L.stream_ptr())
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dres=None, want16=False, x_stride=None):
    rows = dy.shape[0]
    dx = _empty(rows, 768, dtype=torch.float32, device=dy.device) if dres is None else dres
    dx16 = _empty(rows, 768, dtype=torch.bfloat16, device=dy.device) if want16 else None
    dg = _zeros(768, dtype=torch.float32, device=dy.device)
    db = _zeros(*dg.shape, dtype=dg.dtype, device=dg.device)
    cs = _zeros(768, dtype=torch.float32, device=dy.device)
    L.call("vitk_layernorm_bwd", L.ptr(dy), DT[dy.dtype], L.ptr(x), x_stride or 768, L.ptr(gamma), L.ptr(mean),
           L.ptr(rstd), L.ptr(dres), L.ptr(dx), L.ptr(dx16), L.ptr(dg), L.ptr(db), L.ptr(cs), rows, L.stream_ptr())
    return dx, dx16, dg, db, cs


def linear_fwd(x, w, bias, epilogue, engine, x_layout=L.LAYOUT_ROWMAJOR, residual=None, M=None, scratch=None):
    N, K = w.shape
    M = (x.shape[0] if x_layout == L.LAYOUT_ROWMAJOR else x.shape[1]) if M is None else M
    dt = x.dtype
    aux = None
    if epilogue == L.EPI_BIAS_RESIDUAL:
        y = _empty(M, N, dtype=torch.float32, device=x.device)
        aux = residual
    elif epilogue == L.EPI_QKV_SCATTER:
        y = _empty(N // 64, M, 64, dtype=dt, device=x.device)
    else:
        y = _empty(M, N, dtype=dt, device=x.device)
        if epilogue == L.EPI_BIAS_GELU:
            aux = _empty(M, N, dtype=dt, device=x.device)
    if scratch is not None:     # fp32 scratch (zeroed): permits the row-tail split (vitk_linear_fwd_ws)
        L.call("vitk_linear_fwd_ws", L.ptr(x), x_layout, L.ptr(w), L.ptr(bias), L.ptr(y), L.ptr(aux), M, N, K, epilogue,
               DT[dt], engine, L.ptr(scratch), scratch.numel(), L.stream_ptr())
    else:
        L.call("vitk_linear_fwd", L.ptr(x), x_layout, L.ptr(w), L.ptr(bias), L.ptr(y), L.ptr(aux), M, N, K, epilogue,
               DT[dt], engine, L.stream_ptr())
    return (y, aux) if epilogue == L.EPI_BIAS_GELU else y


def linear_dgrad(dy, w, engine, dy_layout=L.LAYOUT_ROWMAJOR, gelu_grad=None, want_colsum=False, scratch=None):
    N, K = w.shape
    M = dy.shape[0] if dy_layout == L.LAYOUT_ROWMAJOR else dy.shape[1]
    dx = _empty(M, K, dtype=dy.dtype, device=dy.device)
    cs = _zeros(K, dtype=torch.float32, device=dy.device) if want_colsum else None
    if scratch is not None:
        L.call("vitk_linear_dgrad_ws", L.ptr(dy), dy_layout, L.ptr(w), L.ptr(dx), L.ptr(gelu_grad), L.ptr(cs), M, N, K,
               DT[dy.dtype], engine, L.ptr(scratch), scratch.numel(), L.stream_ptr())
    else:
        L.call("vitk_linear_dgrad", L.ptr(dy), dy_layout, L.ptr(w), L.ptr(dx), L.ptr(gelu_grad), L.ptr(cs), M, N, K,
               DT[dy.dtype], engine, L.stream_ptr())
    return (dx, cs) if cs is not None else dx


def linear_wgrad(dy, x, N, K, engine, dy_layout=L.LAYOUT_ROWMAJOR):
    M = x.shape[0]
    dw = _zeros(N, K, dtype=torch.float32, device=x.device)
    db = _zeros(N, dtype=torch.float32, device=x.device)
    L.call("vitk_linear_wgrad", L.ptr(dy), dy_layout, L.ptr(x), L.ptr(dw), L.ptr(db), M, N, K, DT[x.dtype], engine,
           L.stream_ptr())
    return dw, db


def attn_fwd(qkv_hm, batch):
    M = batch * 197
    out = _empty(M, 768, dtype=qkv_hm.dtype, device=qkv_hm.device)
    lse = _empty(12, M, dtype=torch.float32, device=qkv_hm.device)
    L.call("vitk_attn_fwd", L.ptr(qkv_hm), L.ptr(out), L.ptr(lse), batch, DT[qkv_hm.dtype], L.stream_ptr())
    return out, lse


def attn_bwd(qkv_hm, out, dout, lse, batch):
    dqkv = _empty(*qkv_hm.shape, dtype=qkv_hm.dtype, device=qkv_hm.device)
    cs = _zeros(2304, dtype=torch.float32, device=qkv_hm.device)
    L.call("vitk_attn_bwd", L.ptr(qkv_hm), L.ptr(out), L.ptr(dout), L.ptr(lse), L.ptr(dqkv), L.ptr(cs), batch,
           DT[qkv_hm.dtype], L.stream_ptr())
    return dqkv, cs


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
