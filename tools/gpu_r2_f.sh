#!/bin/bash
# round 2, call F: split-tail tests, same-process A/B (knob 9 = 1: plain launches), guard bands / repeatability, quick bench
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_bench_shape_parity.py tests/test_guard_bands.py -m gpu -q -x > gpurun_out/f2_pytest_new.log 2>&1; echo "new tests rc=$?"; tail -15 gpurun_out/f2_pytest_new.log
timeout 200 python tools/knob_ab.py 9:0 9:1 --rounds 4 --steps 10 > gpurun_out/f2_ab.log 2>&1; echo "ab rc=$?"; tail -3 gpurun_out/f2_ab.log
timeout 300 python bench.py --no-cpu-baseline --no-eager-baseline --no-extras --no-sustained > gpurun_out/f2_bench.json 2> gpurun_out/f2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/f2_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'clocks',d['clocks'])
print('roofline',d['roofline']['achieved'],d['roofline']['frac'],d['roofline']['gemm_share_of_step'])
PY
tail -14 gpurun_out/f2_bench.err
