#!/bin/bash
mkdir -p gpurun_out
./tools/micro/mma_rate > gpurun_out/d_mma_rate.log 2>&1
python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 > gpurun_out/d_pytest_kernels.log 2>&1
echo "pytest kernels exit $?" > gpurun_out/d_status.log
python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 900 -k "not loss_curve" > gpurun_out/d_pytest_model.log 2>&1
echo "pytest model exit $?" >> gpurun_out/d_status.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
echo "bench exit $?" >> gpurun_out/d_status.log
python tools/step_profile.py > gpurun_out/d_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 700 --csv --log-file gpurun_out/d_launches.csv python tools/step_profile.py > gpurun_out/d_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/d_status.log
cat gpurun_out/d_status.log gpurun_out/d_mma_rate.log; tail -3 gpurun_out/d_pytest_kernels.log gpurun_out/d_pytest_model.log; tail -16 gpurun_out/d_bench.err
