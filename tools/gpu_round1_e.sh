#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -k "linear" > gpurun_out/e_pytest_kernels.log 2>&1
echo "pytest kernels exit $?" > gpurun_out/e_status.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
echo "bench exit $?" >> gpurun_out/e_status.log
python tools/gemm_profile.py > gpurun_out/e_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 14 -c 7 -o gpurun_out/e_gemm python tools/gemm_profile.py > gpurun_out/e_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/e_status.log
cat gpurun_out/e_status.log; tail -3 gpurun_out/e_pytest_kernels.log; tail -16 gpurun_out/e_bench.err
