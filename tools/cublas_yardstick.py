"""Vendor-library yardstick for the GEMM shapes of the bs-64 training step: torch.matmul (cuBLASLt, bf16, fp32 accumulate, NO
fused epilogue work except what cuBLAS does itself) next to this package's tcgen05 kernel WITH its fused epilogue, both
timed the same way: 24 back-to-back launches between two CUDA events, activations rotating through 4 copies (so the
A operand is not L2-resident from the previous launch), weights shared.  Prints TFLOP/s per shape and the FLOP-weighted
totals for one training step (12 layers x the 12 shapes)."""
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
M = int(os.environ.get("M", "12608"))
E = L.ENGINE_TCGEN05
NCOPY, REPS = 4, 24


def rnd(*s, scale=1.0):
    return (torch.randn(*s, device=DEV) * scale).to(bf)


x768 = [rnd(M, 768) for _ in range(NCOPY)]
x3072 = [rnd(M, 3072) for _ in range(NCOPY)]
x2304 = [rnd(M, 2304) for _ in range(NCOPY)]
hm2304 = [K.to_headmajor(t) for t in x2304]
res = [torch.randn(M, 768, device=DEV) for _ in range(NCOPY)]
w_fc1, w_fc2, w_proj, w_qkv = rnd(3072, 768, scale=.05), rnd(768, 3072, scale=.05), rnd(768, 768, scale=.05), rnd(2304, 768, scale=.05)
b3072, b768, b2304 = torch.randn(3072, device=DEV), torch.randn(768, device=DEV), torch.randn(2304, device=DEV)

# the training-mode configuration: a zeroed scratch lets the deep J = 768 GEMMs take the split tail;
# SCRATCH=0 times the plain launches (eval mode)
SCR = None if os.environ.get("SCRATCH") == "0" else torch.zeros(L.load().vitk_gemm_tail_scratch_floats(768), dtype=torch.float32, device=DEV)

_dw = {}


def wgrad(dy, x, N, Kk, layout=L.LAYOUT_ROWMAJOR):
    # accumulates into a preallocated fp32 gradient with db = NULL, as the model driver does (the bias gradients come out of
    # other kernels' epilogues; a non-NULL db would add a stand-alone column-sum pass over dy to the timing)
    if (N, Kk) not in _dw:
        _dw[(N, Kk)] = torch.zeros(N, Kk, device=DEV)
    L.call("vitk_linear_wgrad", L.ptr(dy), layout, L.ptr(x), L.ptr(_dw[(N, Kk)]), None, M, N, Kk, L.BF16, E, L.stream_ptr())


# name, N*K, ours(i), cublas(i)
CASES = [
    ("qkv fwd   12608x2304x768", 2304 * 768, lambda i: K.linear_fwd(x768[i], w_qkv, b2304, L.EPI_QKV_SCATTER, E), lambda i: x768[i] @ w_qkv.t()),
    ("proj fwd  12608x768x768", 768 * 768, lambda i: K.linear_fwd(x768[i], w_proj, b768, L.EPI_BIAS_RESIDUAL, E, residual=res[i]), lambda i: x768[i] @ w_proj.t()),
    ("fc1 fwd   12608x3072x768", 3072 * 768, lambda i: K.linear_fwd(x768[i], w_fc1, b3072, L.EPI_BIAS_GELU, E), lambda i: x768[i] @ w_fc1.t()),
    ("fc2 fwd   12608x768x3072", 768 * 3072, lambda i: K.linear_fwd(x3072[i], w_fc2, b768, L.EPI_BIAS_RESIDUAL, E, residual=res[i], scratch=SCR), lambda i: x3072[i] @ w_fc2.t()),
    ("fc2 dgrad 12608x3072x768", 3072 * 768, lambda i: K.linear_dgrad(x768[i], w_fc2, E, gelu_grad=x3072[(i + 1) % NCOPY]), lambda i: x768[i] @ w_fc2),
    ("fc1 dgrad 12608x768x3072", 768 * 3072, lambda i: K.linear_dgrad(x3072[i], w_fc1, E, scratch=SCR), lambda i: x3072[i] @ w_fc1),
    ("proj dgrad 12608x768x768", 768 * 768, lambda i: K.linear_dgrad(x768[i], w_proj, E), lambda i: x768[i] @ w_proj),
    ("qkv dgrad 12608x768x2304", 768 * 2304, lambda i: K.linear_dgrad(hm2304[i], w_qkv, E, dy_layout=L.LAYOUT_HEADMAJOR, scratch=SCR), lambda i: x2304[i] @ w_qkv),
    ("fc2 wgrad 768x3072x12608", 768 * 3072, lambda i: wgrad(x768[i], x3072[i], 768, 3072), lambda i: x768[i].t() @ x3072[i]),
    ("fc1 wgrad 3072x768x12608", 768 * 3072, lambda i: wgrad(x3072[i], x768[i], 3072, 768), lambda i: x3072[i].t() @ x768[i]),
    ("proj wgrad 768x768x12608", 768 * 768, lambda i: wgrad(x768[i], x768[(i + 1) % NCOPY], 768, 768), lambda i: x768[i].t() @ x768[(i + 1) % NCOPY]),
    ("qkv wgrad 2304x768x12608", 768 * 2304, lambda i: wgrad(hm2304[i], x768[i], 2304, 768, L.LAYOUT_HEADMAJOR), lambda i: x2304[i].t() @ x768[i]),
]


def timeit(fn):
    for i in range(NCOPY):
        fn(i)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for r in range(REPS):
            fn(r % NCOPY)
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / REPS * 1e3)
    return best


if os.environ.get("ONLY"):
    CASES = [c for c in CASES if os.environ["ONLY"] in c[0]]
print("knobs:", os.environ.get("VITK_KNOBS", ""), " split-tail scratch:", SCR is not None)
print(f"{'shape':28s} {'ours (fused epilogue)':>24s} {'cuBLAS (plain GEMM)':>24s}")
tot_o = tot_c = 0.0
for name, nk, ours, cublas in CASES:
    flops = 2.0 * M * nk
    to, tc = timeit(ours), timeit(cublas)
    tot_o += to
    tot_c += tc
    print(f"{name:28s} {to:8.1f} us {flops / to / 1e6:7.0f} TF   {tc:8.1f} us {flops / tc / 1e6:7.0f} TF", flush=True)
step_flops = 12 * sum(2.0 * M * nk for _, nk, _, _ in CASES)
print(f"one step (12 layers): ours {12 * tot_o / 1e3:.3f} ms ({step_flops / (12 * tot_o) / 1e6:.0f} TF), "
      f"cuBLAS {12 * tot_c / 1e3:.3f} ms ({step_flops / (12 * tot_c) / 1e6:.0f} TF) -- cuBLAS figure excludes bias/GELU/residual/"
      f"scatter passes and the fp32 accumulation of wgrad into the gradient buffer")
