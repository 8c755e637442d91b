#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -x -k "attention" > gpurun_out/r_pytest.log 2>&1
echo "pytest attention exit $?" > gpurun_out/r_status.log
tail -4 gpurun_out/r_pytest.log | cut -c1-300
timeout 300 python tools/attn_bench.py > gpurun_out/r_attn.log 2>&1
echo "attn bench exit $?" >> gpurun_out/r_status.log
cat gpurun_out/r_status.log gpurun_out/r_attn.log
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 900 -x -k "not loss_curve" > gpurun_out/r_pytest_model.log 2>&1
echo "pytest model exit $?" >> gpurun_out/r_status.log
tail -3 gpurun_out/r_pytest_model.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
cut -c1-200 gpurun_out/r_bench.json; grep -o '"extra": {[^}]*}' gpurun_out/r_bench.json
