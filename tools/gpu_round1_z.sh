#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/z_scale.log
for c in 16 32; do
  echo "=== gpus 8 NCCL_MAX_CTAS=$c" >> gpurun_out/z_scale.log
  NCCL_MAX_CTAS=$c NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 295$c bench.py --gpus 8 --steps 10 --warmup 3 2>gpurun_out/z_err_$c.log | cut -c1-200 >> gpurun_out/z_scale.log
done
cat gpurun_out/z_scale.log; grep -i -m5 "nvls\|Channel.*via\|nChannels\|Algo" gpurun_out/z_err_16.log | cut -c1-220; grep -ci nvls gpurun_out/z_err_16.log
