#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench1.json 2> gpurun_out/g_bench1.err
echo "bench1 exit $?" > gpurun_out/g_status.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/g_bench2.json 2> gpurun_out/g_bench2.err
echo "bench2 exit $?" >> gpurun_out/g_status.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dp_check_nccl.py > gpurun_out/g_dpcheck.log 2>&1
echo "dpcheck exit $?" >> gpurun_out/g_status.log
cat gpurun_out/g_status.log; cat gpurun_out/g_bench1.json gpurun_out/g_bench2.json | cut -c1-400; tail -5 gpurun_out/g_bench2.err; tail -8 gpurun_out/g_dpcheck.log
