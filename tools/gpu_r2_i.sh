#!/bin/bash
# round 2, call I: single fp32 staging slot under 256-wide pair tiles (5-stage ring) -- kernel tests, A/B, bench with the inference extras
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bench_shape_parity.py tests/test_guard_bands.py tests/test_gpu_kernels.py tests/test_full_size_properties.py -m gpu -q -x > gpurun_out/i_pytest.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/i_pytest.log
timeout 200 python tools/knob_ab.py 9:0 9:1 --rounds 4 --steps 10 > gpurun_out/i_ab.log 2>&1; echo "ab rc=$?"; tail -3 gpurun_out/i_ab.log
timeout 400 python bench.py --no-cpu-baseline --no-eager-baseline > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/i_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'clocks',d['clocks'])
print('roofline',d['roofline']['achieved'],d['roofline']['frac'],d['roofline']['gemm_share_of_step'])
x=d['extra']
print({k:x[k] for k in x if k.startswith('bs') or k.startswith('frozen') or k.startswith('mfu') or k.startswith('eager')})
print(x.get('sustained'))
PY
tail -14 gpurun_out/i_bench.err
