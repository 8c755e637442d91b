#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 > gpurun_out/h_pytest_kernels.log 2>&1
echo "pytest kernels exit $?" > gpurun_out/h_status.log
python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 900 -k "not loss_curve" > gpurun_out/h_pytest_model.log 2>&1
echo "pytest model exit $?" >> gpurun_out/h_status.log
python tools/launch_overhead.py > gpurun_out/h_overhead.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
echo "bench exit $?" >> gpurun_out/h_status.log
python tools/step_profile.py > gpurun_out/h_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 700 --csv --log-file gpurun_out/h_launches.csv python tools/step_profile.py > gpurun_out/h_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/h_status.log
cat gpurun_out/h_status.log gpurun_out/h_overhead.log; tail -3 gpurun_out/h_pytest_kernels.log gpurun_out/h_pytest_model.log | cut -c1-200; tail -16 gpurun_out/h_bench.err
