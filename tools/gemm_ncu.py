"""A handful of tcgen05 GEMM launches for `ncu --set full -k regex:gemm_tc`: qkv forward (scatter epilogue) and fc1
forward (GELU epilogue) at the bs-64 shape, single-CTA and CTA-pair kernels, BLOCK_N 256 (launch order printed)."""
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
w_qkv = (torch.randn(2304, 768, device=DEV) * 0.05).to(bf)
w_fc1 = (torch.randn(3072, 768, device=DEV) * 0.05).to(bf)
b2304, b3072 = torch.randn(2304, device=DEV), torch.randn(3072, device=DEV)
order = []
for cg in (1, 2):
    lib.vitk_debug_set(4, cg)
    lib.vitk_debug_set(2, 256)
    K.linear_fwd(x768, w_qkv, b2304, L.EPI_QKV_SCATTER, E); order.append(f"qkv fwd scatter cg{cg} bn256")
    K.linear_fwd(x768, w_fc1, b3072, L.EPI_BIAS_GELU, E); order.append(f"fc1 fwd gelu cg{cg} bn256")
torch.cuda.synchronize()
print("\n".join(order))
