#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 > gpurun_out/f_pytest_kernels.log 2>&1
echo "pytest kernels exit $?" > gpurun_out/f_status.log
python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 900 -k "not loss_curve" > gpurun_out/f_pytest_model.log 2>&1
echo "pytest model exit $?" >> gpurun_out/f_status.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
echo "bench exit $?" >> gpurun_out/f_status.log
python tools/step_profile.py > gpurun_out/f_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 700 --csv --log-file gpurun_out/f_launches.csv python tools/step_profile.py > gpurun_out/f_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/f_status.log
cat gpurun_out/f_status.log; tail -3 gpurun_out/f_pytest_kernels.log gpurun_out/f_pytest_model.log | cut -c1-200; tail -16 gpurun_out/f_bench.err
