#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x > gpurun_out/p_pytest.log 2>&1
echo "pytest kernels exit $?" > gpurun_out/p_status.log
tail -4 gpurun_out/p_pytest.log | cut -c1-300
if grep -q "pytest kernels exit 0" gpurun_out/p_status.log; then
  timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 900 -x -k "not loss_curve" > gpurun_out/p_pytest_model.log 2>&1
  echo "pytest model exit $?" >> gpurun_out/p_status.log
  tail -3 gpurun_out/p_pytest_model.log | cut -c1-300
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err
  echo "bench exit $?" >> gpurun_out/p_status.log
  VITK_NO_PDL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/p_bench_nopdl.json 2> gpurun_out/p_bench_nopdl.err
  echo "bench nopdl exit $?" >> gpurun_out/p_status.log
  timeout 600 python tools/gemm_bench.py > gpurun_out/p_gemm.log 2>&1
  echo "gemm bench exit $?" >> gpurun_out/p_status.log
else
  grep -E "Error|error|FAILED|assert" gpurun_out/p_pytest.log | head -20 | cut -c1-300
fi
cat gpurun_out/p_status.log gpurun_out/p_gemm.log; cut -c1-200 gpurun_out/p_bench.json; echo; cut -c1-200 gpurun_out/p_bench_nopdl.json; tail -16 gpurun_out/p_bench.err
