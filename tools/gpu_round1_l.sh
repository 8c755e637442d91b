#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x -k "tcgen05 or attention or linear" > gpurun_out/l_pytest.log 2>&1
echo "pytest exit $?" > gpurun_out/l_status.log
tail -5 gpurun_out/l_pytest.log | cut -c1-300
if grep -q "pytest exit 0" gpurun_out/l_status.log; then
  timeout 300 python tools/attn_bench.py > gpurun_out/l_attn.log 2>&1
  echo "attn bench exit $?" >> gpurun_out/l_status.log
  timeout 600 python tools/gemm_bench.py > gpurun_out/l_gemm.log 2>&1
  echo "gemm bench exit $?" >> gpurun_out/l_status.log
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
  echo "bench exit $?" >> gpurun_out/l_status.log
fi
cat gpurun_out/l_status.log gpurun_out/l_attn.log gpurun_out/l_gemm.log; cut -c1-200 gpurun_out/l_bench.json; tail -16 gpurun_out/l_bench.err
