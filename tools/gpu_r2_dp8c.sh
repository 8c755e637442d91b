#!/bin/bash
mkdir -p gpurun_out
VITK_NVLS_PROF=1 VITK_DP_MODE=nvls VITK_NVLS_BCAST_CTAS=32 timeout 300 python bench.py --gpus 8 --steps 20 --warmup 5 --no-sustained --no-eager-baseline --no-extras > gpurun_out/dp8_prof.json 2> gpurun_out/dp8_prof.err; echo "rc=$?"
grep "NVLS step phases" gpurun_out/dp8_prof.err
VITK_NVLS_PROF=1 VITK_DP_MODE=nvls VITK_NVLS_EAGER_MASTERS=1 timeout 300 python bench.py --gpus 8 --steps 20 --warmup 5 --no-sustained --no-eager-baseline --no-extras > gpurun_out/dp8_prof_e.json 2> gpurun_out/dp8_prof_e.err; echo "rc=$?"
grep "NVLS step phases" gpurun_out/dp8_prof_e.err | head -3
python -c "
import json
for n in ('prof','prof_e'):
    d=json.load(open('gpurun_out/dp8_%s.json'%n)); print(n, d['value'], d['ms_per_step'])"
