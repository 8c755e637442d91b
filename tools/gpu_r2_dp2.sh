#!/bin/bash
# 2-GPU: data-parallel parity (gloo / nccl / nvls) and the two collective modes in the bench
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/dp2_topo.txt 2>&1
timeout 900 python -m pytest tests/test_dp_parity.py -x -q > gpurun_out/dp2_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/dp2_pytest.log
for mode in nvls nccl; do
  VITK_DP_MODE=$mode timeout 300 python bench.py --gpus 2 --steps 20 --warmup 5 --no-sustained --no-eager-baseline > gpurun_out/dp2_bench_$mode.json 2> gpurun_out/dp2_bench_$mode.err; echo "bench $mode rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/dp2_bench_$mode.json'))
print('$mode', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['collective'][:60])" 2>&1 | tail -1
done
timeout 200 python bench.py --steps 20 --warmup 5 --no-sustained --no-eager-baseline --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('single', d['value'], d['ms_per_step'])"
