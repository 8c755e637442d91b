#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j_bench1.json 2> gpurun_out/j_bench1.err
echo "bench1 exit $?" > gpurun_out/j_status.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/j_bench2.json 2> gpurun_out/j_bench2.err
echo "bench2 exit $?" >> gpurun_out/j_status.log
cat gpurun_out/j_status.log; cut -c1-300 gpurun_out/j_bench1.json; echo; cut -c1-300 gpurun_out/j_bench2.json; tail -5 gpurun_out/j_bench2.err | cut -c1-300
