#!/bin/bash
# 8-GPU: the two collective modes, then N=1 on the same box
mkdir -p gpurun_out
for mode in nvls nccl; do
  VITK_DP_MODE=$mode timeout 300 python bench.py --gpus 8 --steps 20 --warmup 5 --no-sustained --no-eager-baseline --no-extras > gpurun_out/dp8_bench_$mode.json 2> gpurun_out/dp8_bench_$mode.err; echo "bench $mode rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/dp8_bench_$mode.json'))
print('$mode', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])" 2>&1 | tail -1
done
timeout 200 python bench.py --steps 20 --warmup 5 --no-sustained --no-eager-baseline --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('single', d['value'], d['ms_per_step'], d['e2e']['value'])"
