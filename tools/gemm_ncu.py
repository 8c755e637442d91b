"""A handful of tcgen05 GEMM launches for `ncu --set full -k regex:gemm_tc`: qkv forward (scatter epilogue, K=768),
fc2 forward (residual epilogue, K=3072), fc1 forward (GELU) and fc1 wgrad (stream-K) at the bs-64 shape with the
heuristic's tile choice (launch order printed)."""
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
lib = L.load()
x768 = (torch.randn(M, 768, device=DEV)).to(bf)
x3072 = (torch.randn(M, 3072, device=DEV)).to(bf)
w_qkv = (torch.randn(2304, 768, device=DEV) * 0.05).to(bf)
w_fc1 = (torch.randn(3072, 768, device=DEV) * 0.05).to(bf)
w_fc2 = (torch.randn(768, 3072, device=DEV) * 0.05).to(bf)
b2304, b3072, b768 = torch.randn(2304, device=DEV), torch.randn(3072, device=DEV), torch.randn(768, device=DEV)
res = torch.randn(M, 768, device=DEV)
order = []
for rep in range(2):
    K.linear_fwd(x768, w_qkv, b2304, L.EPI_QKV_SCATTER, E); order.append("qkv fwd scatter")
    K.linear_fwd(x3072, w_fc2, b768, L.EPI_BIAS_RESIDUAL, E, residual=res); order.append("fc2 fwd residual")
    K.linear_fwd(x768, w_fc1, b3072, L.EPI_BIAS_GELU, E); order.append("fc1 fwd gelu")
    K.linear_wgrad(x3072, x768, 3072, 768, E); order.append("fc1 wgrad")
torch.cuda.synchronize()
print("\n".join(order))
