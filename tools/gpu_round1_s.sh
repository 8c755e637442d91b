#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/s_attn.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -x -k "attention" > gpurun_out/s_pytest.log 2>&1
echo "pytest exit $?" > gpurun_out/s_status.log
tail -3 gpurun_out/s_pytest.log | cut -c1-300
VITK_KNOBS="" timeout 300 python tools/attn_bench.py >> gpurun_out/s_attn.log 2>&1
cat gpurun_out/s_status.log gpurun_out/s_attn.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err
cut -c1-200 gpurun_out/s_bench.json; tail -15 gpurun_out/s_bench.err;  grep -o '"extra": {[^}]*}' gpurun_out/s_bench.json
