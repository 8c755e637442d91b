#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dp_check_nccl.py > gpurun_out/w_dpcheck.log 2>&1
echo "dpcheck exit $?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/w_bench2.json 2> gpurun_out/w_bench2.err
echo "bench2 exit $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 3 > gpurun_out/w_ref2.json 2>/dev/null
echo "ref2 exit $?"
tail -3 gpurun_out/w_dpcheck.log | cut -c1-200; cut -c1-200 gpurun_out/w_bench2.json; cut -c1-120 gpurun_out/w_ref2.json
