"""Fixed per-launch cost of the tcgen05 GEMM kernel: back-to-back launches of tiny / small / proj-sized GEMMs,
CUDA-event timed (total / launches), versus torch.matmul (cuBLAS) on the same shapes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from vit_spoof_detection_pda_b200 import _lib as L  # noqa: E402
import kernels_api as K  # noqa: E402

DEV = torch.device("cuda:0")
bf = torch.bfloat16


def bench(fn, n=200):
    for _ in range(10):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


for (M, N, Kd) in [(128, 192, 64), (128, 768, 768), (2048, 768, 768), (12608, 768, 768), (12608, 2304, 768), (12608, 768, 3072)]:
    x = torch.randn(M, Kd, device=DEV).to(bf)
    w = (torch.randn(N, Kd, device=DEV) * 0.05).to(bf)
    b = torch.randn(N, device=DEV)
    y = torch.empty(M, N, dtype=bf, device=DEV)
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream

    def ours():
        lib.vitk_linear_fwd(x.data_ptr(), 0, w.data_ptr(), b.data_ptr(), y.data_ptr(), None, M, N, Kd, L.EPI_BIAS, L.BF16,
                            L.ENGINE_TCGEN05, st)

    t_ours = bench(ours)
    t_cublas = bench(lambda: torch.matmul(x, w.t()))
    ideal = 2.0 * M * N * Kd / 1351.4e12 * 1e6
    print(f"[{M}x{N}x{Kd}] ours {t_ours:7.2f} us  cuBLAS {t_cublas:7.2f} us  ideal@1351TF {ideal:6.2f} us", flush=True)
