#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/attn_bench.py > gpurun_out/ab_attn.log 2>&1; echo "attn exit $?" > gpurun_out/ab_status.log
timeout 300 python tools/knob_ab.py 12:0 12:1 --rounds 4 --steps 10 > gpurun_out/ab_knob12.log 2>&1; echo "knob exit $?" >> gpurun_out/ab_status.log
timeout 300 python tools/cublas_yardstick.py > gpurun_out/ab_cublas.log 2>&1; echo "cublas exit $?" >> gpurun_out/ab_status.log
cat gpurun_out/ab_status.log gpurun_out/ab_attn.log gpurun_out/ab_knob12.log; tail -n 16 gpurun_out/ab_cublas.log
