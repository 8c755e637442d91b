#!/bin/bash
# 8-GPU data-parallel bench (BASELINE configs[4]) after the round's last kernel changes
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/al_bench8.json 2> gpurun_out/al_bench8.err
echo "bench8 exit $?" > gpurun_out/al_status.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/al_bench1.json 2> gpurun_out/al_bench1.err
echo "bench1 exit $?" >> gpurun_out/al_status.log
cat gpurun_out/al_status.log; cut -c1-200 gpurun_out/al_bench8.json; echo; cut -c1-200 gpurun_out/al_bench1.json
