#!/bin/bash
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" VITK_NVLS_PROF=1 timeout 240 python bench.py --gpus 8 --steps 20 --warmup 5 --no-sustained --no-eager-baseline --no-extras > gpurun_out/dp8_$name.json 2> gpurun_out/dp8_$name.err; echo "bench $name rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/dp8_$name.json'))
print('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])" 2>&1 | tail -1
  grep "rank 0. NVLS step phases" gpurun_out/dp8_$name.err
}
run d4 VITK_DP_MODE=nvls
run d8 VITK_DP_MODE=nvls VITK_NVLS_DOMAINS=8
run d4b VITK_DP_MODE=nvls
timeout 200 python bench.py --steps 20 --warmup 5 --no-sustained --no-eager-baseline --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('single', d['value'], d['ms_per_step'], d['e2e']['value'])"
