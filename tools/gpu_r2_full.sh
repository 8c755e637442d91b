#!/bin/bash
# full GPU test suite (with a global timeout) + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/f_pytest.log
timeout 600 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/f_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'clocks',d['clocks'])
print('roofline',d['roofline']['achieved'],d['roofline']['frac'])
x=d['extra']
print({k:x[k] for k in x if k.startswith('bs') or k.startswith('frozen') or k.startswith('mfu') or k.startswith('speedup')})
print(x.get('sustained')); print(x.get('hbm_kernels'))
PY
tail -16 gpurun_out/f_bench.err
