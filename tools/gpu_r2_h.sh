#!/bin/bash
# round 2, call H (2 GPUs): data-parallel parity (NCCL bucketed path and the NVLS fused optimizer step) and a 2-GPU bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_parity.py -x -q > gpurun_out/h_pytest_dp.log 2>&1; echo "dp pytest rc=$?"; tail -4 gpurun_out/h_pytest_dp.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-sustained --no-extras > gpurun_out/h_bench2.json 2> gpurun_out/h_bench2.err; echo "bench2 rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/h_bench2.json'))
print('N=2', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['collective'][:60], d['clocks'])"
tail -3 gpurun_out/h_bench2.err
