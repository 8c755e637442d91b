#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err
echo "bench exit $?" > gpurun_out/t_status.log
python tools/step_profile.py > gpurun_out/t_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 330 --csv --log-file gpurun_out/t_launches.csv python tools/step_profile.py > gpurun_out/t_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/t_status.log
cat gpurun_out/t_status.log; cut -c1-200 gpurun_out/t_bench.json; tail -15 gpurun_out/t_bench.err
