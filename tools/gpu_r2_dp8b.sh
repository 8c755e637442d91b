#!/bin/bash
# 8-GPU A/B on one box: NVLS fused step with lazy vs eager fp32-master broadcast, then N=1
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --gpus 8 --steps 20 --warmup 5 --no-sustained --no-eager-baseline --no-extras > gpurun_out/dp8_$name.json 2> gpurun_out/dp8_$name.err; echo "bench $name rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/dp8_$name.json'))
print('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])" 2>&1 | tail -1
}
run lazy VITK_DP_MODE=nvls
run eager VITK_DP_MODE=nvls VITK_NVLS_EAGER_MASTERS=1
run lazy32 VITK_DP_MODE=nvls VITK_NVLS_BCAST_CTAS=32
run lazy_b VITK_DP_MODE=nvls
timeout 200 python bench.py --steps 20 --warmup 5 --no-sustained --no-eager-baseline --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('single', d['value'], d['ms_per_step'], d['e2e']['value'])"
