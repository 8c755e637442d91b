"""GPU probe for the tcgen05 GEMM engine: runs every operand-layout / epilogue combination against a
torch fp32 reference and prints the error of each (no asserts, never stops at the first failure).
Also sweeps the debug descriptor variants so one GPU call answers which encoding is right."""
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from vit_spoof_detection_pda_b200 import _lib as L  # noqa: E402
import kernels_api as K  # noqa: E402

DEV = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False


def randn(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, generator=g, device=DEV) * scale


def report(name, fn):
    try:
        err = fn()
        torch.cuda.synchronize()
        print(f"{name:60s} rel_err {err:.3e} {'OK' if err < 2e-2 else 'BAD'}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"{name:60s} EXC {e!r}", flush=True)
        traceback.print_exc()


def main():
    lib = L.load()
    E = L.ENGINE_TCGEN05
    bf = torch.bfloat16
    for (M, N, Kd) in [(128, 256, 64), (128, 256, 768), (197, 768, 768), (1576, 2304, 768), (1000, 768, 3072), (12608, 3072, 768)]:
        x = randn(M, Kd, seed=1).to(bf)
        w = randn(N, Kd, seed=2, scale=0.05).to(bf)
        b = randn(N, seed=3, scale=0.5)
        res = randn(M, N, seed=4)
        ref = x.float() @ w.float().t() + b
        tag = f"[{M}x{N}x{Kd}]"
        report(f"fwd bias {tag}", lambda: K.rel_err(K.linear_fwd(x, w, b, L.EPI_BIAS, E).float(), ref))
        report(f"fwd gelu {tag}", lambda: K.rel_err(K.linear_fwd(x, w, b, L.EPI_BIAS_GELU, E)[0].float(), torch.nn.functional.gelu(ref)))
        report(f"fwd residual {tag}", lambda: K.rel_err(K.linear_fwd(x, w, b, L.EPI_BIAS_RESIDUAL, E, residual=res), ref + res))
        report(f"fwd scatter {tag}", lambda: K.rel_err(K.from_headmajor(K.linear_fwd(x, w, b, L.EPI_QKV_SCATTER, E)).float(), ref))
        report(f"fwd bias BN128 {tag}", lambda: (lib.vitk_debug_set(2, 128), K.rel_err(K.linear_fwd(x, w, b, L.EPI_BIAS, E).float(), ref), lib.vitk_debug_set(2, 0))[1])
        dy = randn(M, N, seed=5).to(bf)
        refd = dy.float() @ w.float()
        refw = dy.float().t() @ x.float()
        for variant in (0, 1):
            lib.vitk_debug_set(0, variant)
            report(f"dgrad (B MN-major) v{variant} {tag}", lambda: K.rel_err(K.linear_dgrad(dy, w, E).float(), refd))
            report(f"dgrad headmajor A v{variant} {tag}", lambda: K.rel_err(K.linear_dgrad(K.to_headmajor(dy), w, E, dy_layout=L.LAYOUT_HEADMAJOR).float(), refd))
            report(f"wgrad (A,B MN-major) v{variant} {tag}", lambda: K.rel_err(K.linear_wgrad(dy, x, N, Kd, E)[0], refw))
            report(f"wgrad headmajor A v{variant} {tag}", lambda: K.rel_err(K.linear_wgrad(K.to_headmajor(dy), x, N, Kd, E, dy_layout=L.LAYOUT_HEADMAJOR)[0], refw))
            lib.vitk_debug_set(1, 1)
            report(f"wgrad nosplit v{variant} {tag}", lambda: K.rel_err(K.linear_wgrad(dy, x, N, Kd, E)[0], refw))
            lib.vitk_debug_set(1, 0)
        lib.vitk_debug_set(0, 0)
    # timing of the big forward shapes (bs 64)
    M = 12608
    for (N, Kd) in [(2304, 768), (768, 768), (3072, 768), (768, 3072)]:
        x = randn(M, Kd, seed=1).to(bf)
        w = randn(N, Kd, seed=2, scale=0.05).to(bf)
        b = randn(N, seed=3)
        for _ in range(3):
            K.linear_fwd(x, w, b, L.EPI_BIAS, E)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            K.linear_fwd(x, w, b, L.EPI_BIAS, E)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        print(f"time fwd [{M}x{N}x{Kd}] {ms*1e3:.1f} us  {2*M*N*Kd/ms/1e9:.1f} TFLOP/s", flush=True)
        for _ in range(3):
            torch.matmul(x, w.t())
        s.record()
        for _ in range(20):
            torch.matmul(x, w.t())
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        print(f"  cuBLAS same shape      {ms*1e3:.1f} us  {2*M*N*Kd/ms/1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
