"""Thin torch-tensor wrappers over the C-ABI used by the GPU parity tests (calls go through ctypes,
exactly as the product's Python host code does)."""
import torch

from vit_spoof_detection_pda_b200 import _lib as L

DT = {torch.float32: L.F32, torch.bfloat16: L.BF16}


def to_headmajor(x):  # [M, C] -> [C/64, M, 64]
    M, Cc = x.shape
    return x.reshape(M, Cc // 64, 64).permute(1, 0, 2).contiguous()


def from_headmajor(x):  # [C/64, M, 64] -> [M, C]
    nb, M, _ = x.shape
    return x.permute(1, 0, 2).reshape(M, nb * 64).contiguous()


def layernorm_fwd(x, gamma, beta, eps, out_dtype, x_stride=None, rows=None):
    rows = x.shape[0] if rows is None else rows
    y = torch.empty(rows, 768, dtype=out_dtype, device=x.device)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    L.call("vitk_layernorm_fwd", L.ptr(x), x_stride or 768, L.ptr(gamma), L.ptr(beta), L.ptr(y), DT[out_dtype],
           L.ptr(mean), L.ptr(rstd), rows, eps, L.stream_ptr())
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dres=None, want16=False, x_stride=None):
    rows = dy.shape[0]
    dx = torch.empty(rows, 768, dtype=torch.float32, device=dy.device) if dres is None else dres
    dx16 = torch.empty(rows, 768, dtype=torch.bfloat16, device=dy.device) if want16 else None
    dg = torch.zeros(768, dtype=torch.float32, device=dy.device)
    db = torch.zeros_like(dg)
    cs = torch.zeros(768, dtype=torch.float32, device=dy.device)
    L.call("vitk_layernorm_bwd", L.ptr(dy), DT[dy.dtype], L.ptr(x), x_stride or 768, L.ptr(gamma), L.ptr(mean),
           L.ptr(rstd), L.ptr(dres), L.ptr(dx), L.ptr(dx16), L.ptr(dg), L.ptr(db), L.ptr(cs), rows, L.stream_ptr())
    return dx, dx16, dg, db, cs


def linear_fwd(x, w, bias, epilogue, engine, x_layout=L.LAYOUT_ROWMAJOR, residual=None, M=None):
    N, K = w.shape
    M = (x.shape[0] if x_layout == L.LAYOUT_ROWMAJOR else x.shape[1]) if M is None else M
    dt = x.dtype
    aux = None
    if epilogue == L.EPI_BIAS_RESIDUAL:
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        aux = residual
    elif epilogue == L.EPI_QKV_SCATTER:
        y = torch.empty(N // 64, M, 64, dtype=dt, device=x.device)
    else:
        y = torch.empty(M, N, dtype=dt, device=x.device)
        if epilogue == L.EPI_BIAS_GELU:
            aux = torch.empty(M, N, dtype=dt, device=x.device)
    L.call("vitk_linear_fwd", L.ptr(x), x_layout, L.ptr(w), L.ptr(bias), L.ptr(y), L.ptr(aux), M, N, K, epilogue,
           DT[dt], engine, L.stream_ptr())
    return (y, aux) if epilogue == L.EPI_BIAS_GELU else y


def linear_dgrad(dy, w, engine, dy_layout=L.LAYOUT_ROWMAJOR, gelu_grad=None, want_colsum=False):
    N, K = w.shape
    M = dy.shape[0] if dy_layout == L.LAYOUT_ROWMAJOR else dy.shape[1]
    dx = torch.empty(M, K, dtype=dy.dtype, device=dy.device)
    cs = torch.zeros(K, dtype=torch.float32, device=dy.device) if want_colsum else None
    L.call("vitk_linear_dgrad", L.ptr(dy), dy_layout, L.ptr(w), L.ptr(dx), L.ptr(gelu_grad), L.ptr(cs), M, N, K,
           DT[dy.dtype], engine, L.stream_ptr())
    return (dx, cs) if cs is not None else dx


def linear_wgrad(dy, x, N, K, engine, dy_layout=L.LAYOUT_ROWMAJOR):
    M = x.shape[0]
    dw = torch.zeros(N, K, dtype=torch.float32, device=x.device)
    db = torch.zeros(N, dtype=torch.float32, device=x.device)
    L.call("vitk_linear_wgrad", L.ptr(dy), dy_layout, L.ptr(x), L.ptr(dw), L.ptr(db), M, N, K, DT[x.dtype], engine,
           L.stream_ptr())
    return dw, db


def attn_fwd(qkv_hm, batch):
    M = batch * 197
    out = torch.empty(M, 768, dtype=qkv_hm.dtype, device=qkv_hm.device)
    lse = torch.empty(12, M, dtype=torch.float32, device=qkv_hm.device)
    L.call("vitk_attn_fwd", L.ptr(qkv_hm), L.ptr(out), L.ptr(lse), batch, DT[qkv_hm.dtype], L.stream_ptr())
    return out, lse


def attn_bwd(qkv_hm, out, dout, lse, batch):
    dqkv = torch.empty_like(qkv_hm)
    cs = torch.zeros(2304, dtype=torch.float32, device=qkv_hm.device)
    L.call("vitk_attn_bwd", L.ptr(qkv_hm), L.ptr(out), L.ptr(dout), L.ptr(lse), L.ptr(dqkv), L.ptr(cs), batch,
           DT[qkv_hm.dtype], L.stream_ptr())
    return dqkv, cs


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
