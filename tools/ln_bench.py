"""LayerNorm forward at the bs-64 shape (12608 x 768 fp32 -> bf16), L2 flushed, for a few persistent-grid sizes."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vit_spoof_detection_pda_b200 import _lib as L
dev = torch.device("cuda:0")
M = 12608
x = torch.randn(M, 768, device=dev); g = torch.randn(768, device=dev); b = torch.randn(768, device=dev)
y = torch.empty(M, 768, dtype=torch.bfloat16, device=dev); mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = L.load(); st = torch.cuda.current_stream().cuda_stream
for k in (1, 2, 3, 4, 8):
    lib.vitk_debug_set(9, k)
    tot = 0.0
    for it in range(13):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); L.call("vitk_layernorm_fwd", x.data_ptr(), 768, g.data_ptr(), b.data_ptr(), y.data_ptr(), L.BF16, mean.data_ptr(), rstd.data_ptr(), M, 1e-6, st); e.record(); torch.cuda.synchronize()
        if it >= 3: tot += s.elapsed_time(e)
    print(f"CTAs/SM {k}: {tot / 10 * 1e3:6.1f} us  ({M * 768 * 6 / (tot / 10 * 1e-3) / 1e9:6.0f} GB/s algorithmic)")
