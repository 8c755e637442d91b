"""Runs the representative tcgen05 GEMMs of the bs-64 step once each (after a warm-up pass) -- wrapped by
`ncu --set full -k regex:gemm_tc` to get L2/DRAM/tensor-pipe counters and stall reasons per shape."""
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
M = 12608
E = L.ENGINE_TCGEN05


def rnd(*s, scale=1.0):
    return (torch.randn(*s, device=DEV) * scale).to(bf)


x768, x3072 = rnd(M, 768), rnd(M, 3072)
w_fc1, w_fc2, w_proj = rnd(3072, 768, scale=0.05), rnd(768, 3072, scale=0.05), rnd(768, 768, scale=0.05)
b3072, b768 = torch.randn(3072, device=DEV), torch.randn(768, device=DEV)
res = torch.randn(M, 768, device=DEV)
u = rnd(M, 3072)


def run_all():
    K.linear_fwd(x768, w_fc1, b3072, L.EPI_BIAS_GELU, E)                  # fc1 fwd + GELU      (K-major, epi1)
    K.linear_fwd(x3072, w_fc2, b768, L.EPI_BIAS_RESIDUAL, E, residual=res)  # fc2 fwd + residual  (epi2, K=3072)
    K.linear_fwd(x768, w_proj, b768, L.EPI_BIAS_RESIDUAL, E, residual=res)  # proj fwd + residual (epi2, K=768)
    K.linear_dgrad(x768, w_fc2, E, gelu_grad=u)                              # fc2 dgrad x GELU'   (B MN-major, epi4)
    K.linear_dgrad(x3072, w_fc1, E)                                       # fc1 dgrad           (epi0, K=3072)
    K.linear_wgrad(x768, x3072, 768, 3072, E)                             # fc2 wgrad           (both MN-major, epi5)


for _ in range(2):
    run_all()
torch.cuda.synchronize()
run_all()
torch.cuda.synchronize()
print("done")
