#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dp_parity.py -x -q -k nvls > gpurun_out/dp2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/dp2_pytest.log
VITK_NVLS_PROF=1 VITK_DP_MODE=nvls timeout 200 python bench.py --gpus 2 --steps 20 --warmup 5 --no-sustained --no-eager-baseline --no-extras > gpurun_out/dp2_bench_nvls.json 2> gpurun_out/dp2_bench_nvls.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/dp2_bench_nvls.json'))
print('nvls', d['value'], d['ms_per_step'], d['e2e'])"
grep "NVLS step phases" gpurun_out/dp2_bench_nvls.err
