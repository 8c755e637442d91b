#!/bin/bash
# 2-GPU data-parallel parity + bench (after sliced split-K / pre-zero changes)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dp_check_nccl.py > gpurun_out/aj_dpcheck.log 2>&1
echo "dpcheck exit $?" > gpurun_out/aj_status.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/aj_bench2.json 2> gpurun_out/aj_bench2.err
echo "bench2 exit $?" >> gpurun_out/aj_status.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/aj_bench1.json 2> gpurun_out/aj_bench1.err
echo "bench1 exit $?" >> gpurun_out/aj_status.log
cat gpurun_out/aj_status.log; tail -n 3 gpurun_out/aj_dpcheck.log | cut -c1-200; cut -c1-200 gpurun_out/aj_bench2.json; echo; cut -c1-200 gpurun_out/aj_bench1.json
