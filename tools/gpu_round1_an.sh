#!/bin/bash
# last check of the round: the default NCCL path after the bucketer refactor (parity), then the opt-in 16-bit all-reduce
mkdir -p gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dp_check_nccl.py > gpurun_out/an_dpcheck.log 2>&1
echo "dpcheck exit $?" > gpurun_out/an_status.log
VITK_GRAD_COMM=bf16 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/an_bench2_bf16.json 2> gpurun_out/an_bench2_bf16.err
echo "bench2 bf16-comm exit $?" >> gpurun_out/an_status.log
cat gpurun_out/an_status.log; tail -n 2 gpurun_out/an_dpcheck.log | cut -c1-220; cut -c1-180 gpurun_out/an_bench2_bf16.json
