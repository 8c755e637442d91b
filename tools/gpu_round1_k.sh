#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -k "attention" > gpurun_out/k_pytest_attn.log 2>&1
echo "pytest attention exit $?" > gpurun_out/k_status.log
timeout 300 python tools/attn_bench.py > gpurun_out/k_attn.log 2>&1
echo "attn bench exit $?" >> gpurun_out/k_status.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err
echo "bench exit $?" >> gpurun_out/k_status.log
cat gpurun_out/k_status.log gpurun_out/k_attn.log; tail -5 gpurun_out/k_pytest_attn.log | cut -c1-300; cut -c1-200 gpurun_out/k_bench.json
