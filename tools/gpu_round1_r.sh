#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -x -k "attention" > gpurun_out/r_pytest.log 2>&1
echo "pytest attention exit $?" > gpurun_out/r_status.log
tail -12 gpurun_out/r_pytest.log | cut -c1-300
timeout 300 python tools/attn_bench.py > gpurun_out/r_attn.log 2>&1
echo "attn bench exit $?" >> gpurun_out/r_status.log
cat gpurun_out/r_status.log gpurun_out/r_attn.log
