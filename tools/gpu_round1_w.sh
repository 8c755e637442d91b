#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/w_scale.log
run2() {
  echo "=== $1" >> gpurun_out/w_scale.log
  env $1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>/dev/null | cut -c1-160 >> gpurun_out/w_scale.log
}
run2 "VITK_BUCKET_MB=50"
run2 "VITK_BUCKET_MB=25"
run2 "VITK_BUCKET_MB=25 NCCL_MAX_CTAS=4"
run2 "VITK_BUCKET_MB=100"
cat gpurun_out/w_scale.log
