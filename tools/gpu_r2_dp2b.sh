#!/bin/bash
mkdir -p gpurun_out
for mode in nvls nccl; do
  VITK_DP_MODE=$mode timeout 300 python bench.py --gpus 2 --steps 20 --warmup 5 --no-sustained --no-eager-baseline > gpurun_out/dp2_bench_$mode.json 2> gpurun_out/dp2_bench_$mode.err; echo "bench $mode rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/dp2_bench_$mode.json'))
print('$mode', d['value'], d['ms_per_step'], d['e2e'], d['config']['collective'][:60])" 2>&1 | tail -1
done
