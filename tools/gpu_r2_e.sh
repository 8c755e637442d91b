#!/bin/bash
# round 2, call E: new tests (row-tail split, guard bands, repeatability), same-process A/B of the split, rest of the suite, quick bench
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_bench_shape_parity.py tests/test_guard_bands.py -m gpu -q > gpurun_out/e_pytest_new.log 2>&1; echo "new tests rc=$?"; tail -15 gpurun_out/e_pytest_new.log
timeout 200 python tools/knob_ab.py 9:0 9:1 --rounds 4 --steps 10 > gpurun_out/e_ab.log 2>&1; echo "ab rc=$?"; tail -3 gpurun_out/e_ab.log
timeout 600 python -m pytest tests -m gpu -x -q --ignore=tests/test_bench_shape_parity.py --ignore=tests/test_guard_bands.py > gpurun_out/e_pytest_rest.log 2>&1; echo "rest rc=$?"; tail -4 gpurun_out/e_pytest_rest.log
timeout 400 python bench.py --no-cpu-baseline --no-eager-baseline > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/e_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'clocks',d['clocks'])
print('roofline',d['roofline']['achieved'],d['roofline']['frac'],d['roofline']['gemm_share_of_step'])
x=d['extra']
print({k:x[k] for k in x if k.startswith('bs') or k.startswith('frozen') or k.startswith('mfu') or k.startswith('eager')})
print(x.get('sustained'))
PY
tail -16 gpurun_out/e_bench.err
