#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x -k "tcgen05 or linear" > gpurun_out/n_pytest.log 2>&1
echo "pytest exit $?" > gpurun_out/n_status.log
tail -5 gpurun_out/n_pytest.log | cut -c1-300
if grep -q "pytest exit 0" gpurun_out/n_status.log; then
  timeout 600 python tools/gemm_bench.py > gpurun_out/n_gemm.log 2>&1
  echo "gemm bench exit $?" >> gpurun_out/n_status.log
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err
  echo "bench exit $?" >> gpurun_out/n_status.log
  timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 900 -x -k "not loss_curve" > gpurun_out/n_pytest_model.log 2>&1
  echo "pytest model exit $?" >> gpurun_out/n_status.log
  tail -3 gpurun_out/n_pytest_model.log | cut -c1-300
else
  grep -E "Error|error|FAILED|assert" gpurun_out/n_pytest.log | head -20 | cut -c1-300
fi
cat gpurun_out/n_status.log gpurun_out/n_gemm.log; cut -c1-200 gpurun_out/n_bench.json; tail -16 gpurun_out/n_bench.err
