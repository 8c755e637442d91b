#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/c_pytest.log 2>&1
echo "pytest exit $?" > gpurun_out/c_status.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
echo "bench exit $?" >> gpurun_out/c_status.log
python tools/step_profile.py > gpurun_out/c_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 700 --csv --log-file gpurun_out/c_launches.csv python tools/step_profile.py > gpurun_out/c_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/c_status.log
cat gpurun_out/c_status.log; tail -5 gpurun_out/c_pytest.log; cat gpurun_out/c_bench.json; tail -16 gpurun_out/c_bench.err
