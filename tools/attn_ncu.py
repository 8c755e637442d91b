"""One forward and one backward launch of the bf16 attention kernels at B=64 (768 items) for `ncu --set full`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from vit_spoof_detection_pda_b200 import _lib as L  # noqa: E402
import kernels_api as K  # noqa: E402

DEV = torch.device("cuda:0")
B = 64
M = B * 197
bf = torch.bfloat16
lib = L.load()
qkv = (torch.randn(36, M, 64, device=DEV)).to(bf)
dout = torch.randn(M, 768, device=DEV).to(bf)
for _ in range(2):
    out, lse = K.attn_fwd(qkv, B)
    dqkv, _ = K.attn_bwd(qkv, out, dout, lse, B)
torch.cuda.synchronize()
print("done")
