"""Times the bf16 attention kernels at the bench shape (B=64: 768 (b,h) problems) in isolation."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from vit_spoof_detection_pda_b200 import _lib as L  # noqa: E402
import kernels_api as K  # noqa: E402

DEV = torch.device("cuda:0")
B = int(os.environ.get("B", "64"))
M = B * 197
bf = torch.bfloat16
lib = L.load()
qkv = (torch.randn(36, M, 64, device=DEV)).to(bf)
dout = torch.randn(M, 768, device=DEV).to(bf)
out, lse = K.attn_fwd(qkv, B)
dqkv = torch.empty_like(qkv)
cs = torch.zeros(2304, device=DEV)
st = torch.cuda.current_stream().cuda_stream
# a buffer larger than L2 to flush between timed launches
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / n * 1e3


for kv in os.environ.get('VITK_KNOBS', '').split(','):
    if kv:
        lib.vitk_debug_set(int(kv.split(':')[0]), int(kv.split(':')[1]))
print('knobs:', os.environ.get('VITK_KNOBS', ''))
for variant in ((0,) if os.environ.get('VITK_KNOBS') else (0, 1)):   # 0: tcgen05 kernels, 1: persistent 13-warp mma.sync kernels
    lib.vitk_debug_set(3, variant)
    fwd = timeit(lambda: lib.vitk_attn_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, L.BF16, st))
    print(f"variant {variant} attn fwd  B={B}: {fwd:7.1f} us  ({B*12*4*197*197*64/fwd/1e6:.1f} TFLOP/s algorithmic)")
    t = timeit(lambda: lib.vitk_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
                                         None, B, L.BF16, st))
    print(f"variant {variant} attn bwd  B={B}: {t:7.1f} us  ({B*12*10*197*197*64/t/1e6:.1f} TFLOP/s algorithmic 10*N^2*d)")
    t = timeit(lambda: lib.vitk_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
                                         cs.data_ptr(), B, L.BF16, st))
    print(f"variant {variant} attn bwd + full qkv bias gradient (q, k, v sections; the model driver asks for q only): {t:7.1f} us")
lib.vitk_debug_set(3, 0)
