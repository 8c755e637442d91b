"""Times every tcgen05 GEMM shape of the bs-64 training step in isolation (L2 flushed between launches), for the
single-CTA (cta_group::1) and CTA-pair (cta_group::2) kernels and each forced BLOCK_N, next to the heuristic's pick."""
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
lib = L.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def rnd(*s, scale=1.0):
    return (torch.randn(*s, device=DEV) * scale).to(bf)


x768, x3072, x2304 = rnd(M, 768), rnd(M, 3072), rnd(M, 2304)
w_fc1, w_fc2, w_proj, w_qkv = rnd(3072, 768, scale=.05), rnd(768, 3072, scale=.05), rnd(768, 768, scale=.05), rnd(2304, 768, scale=.05)
b3072, b768, b2304 = torch.randn(3072, device=DEV), torch.randn(768, device=DEV), torch.randn(2304, device=DEV)
res = torch.randn(M, 768, device=DEV)
u = rnd(M, 3072)
hm2304 = K.to_headmajor(x2304)

CASES = [
    ("qkv fwd   scatter  12608x2304x768", 2304 * 768, lambda: K.linear_fwd(x768, w_qkv, b2304, L.EPI_QKV_SCATTER, E)),
    ("proj fwd  residual 12608x768x768", 768 * 768, lambda: K.linear_fwd(x768, w_proj, b768, L.EPI_BIAS_RESIDUAL, E, residual=res)),
    ("fc1 fwd   gelu     12608x3072x768", 3072 * 768, lambda: K.linear_fwd(x768, w_fc1, b3072, L.EPI_BIAS_GELU, E)),
    ("fc2 fwd   residual 12608x768x3072", 768 * 3072, lambda: K.linear_fwd(x3072, w_fc2, b768, L.EPI_BIAS_RESIDUAL, E, residual=res)),
    ("fc2 dgrad gelu'    12608x3072x768", 3072 * 768, lambda: K.linear_dgrad(x768, w_fc2, E, gelu_grad=u)),
    ("fc1 dgrad store    12608x768x3072", 768 * 3072, lambda: K.linear_dgrad(x3072, w_fc1, E)),
    ("proj dgrad store   12608x768x768", 768 * 768, lambda: K.linear_dgrad(x768, w_proj, E)),
    ("qkv dgrad hm       12608x768x2304", 768 * 2304, lambda: K.linear_dgrad(hm2304, w_qkv, E, dy_layout=L.LAYOUT_HEADMAJOR)),
    ("fc2 wgrad          768x3072x12608", 768 * 3072, lambda: K.linear_wgrad(x768, x3072, 768, 3072, E)),
    ("fc1 wgrad          3072x768x12608", 768 * 3072, lambda: K.linear_wgrad(x3072, x768, 3072, 768, E)),
    ("proj wgrad         768x768x12608", 768 * 768, lambda: K.linear_wgrad(x768, x768, 768, 768, E)),
    ("qkv wgrad hm       2304x768x12608", 768 * 2304, lambda: K.linear_wgrad(hm2304, x768, 2304, 768, E, dy_layout=L.LAYOUT_HEADMAJOR)),
]


def timeit(fn, n=6):
    for _ in range(2):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / n * 1e3


for kv in os.environ.get("VITK_KNOBS", "").split(","):   # e.g. VITK_KNOBS=7:1,5:1  -> vitk_debug_set(7, 1); vitk_debug_set(5, 1)
    if kv:
        lib.vitk_debug_set(int(kv.split(":")[0]), int(kv.split(":")[1]))
print("knobs:", os.environ.get("VITK_KNOBS", ""))
COMBOS = [(cg, bn) for cg in (1, 2) for bn in (128, 192, 256)] if os.environ.get("FULL") else [(2, 192), (2, 256)]
print(f"{'case':38s} {'auto':>8s} | " + " ".join(f"cg{cg}/bn{bn:3d}" for cg, bn in COMBOS))
for name, nk, fn in CASES:
    flops = 2.0 * M * nk
    lib.vitk_debug_set(2, 0); lib.vitk_debug_set(4, 0)
    t_auto = timeit(fn)
    cells = []
    for cg, bn in COMBOS:
        lib.vitk_debug_set(2, bn); lib.vitk_debug_set(4, cg)
        try:
            t = timeit(fn)
            cells.append(f"{t:6.1f}us ")
        except Exception:  # noqa: BLE001
            cells.append("    n/a  ")
    lib.vitk_debug_set(2, 0); lib.vitk_debug_set(4, 0)
    print(f"{name:38s} {t_auto:6.1f}us {flops / t_auto / 1e6:6.0f}TF | " + " ".join(cells), flush=True)
