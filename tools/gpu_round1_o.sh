#!/bin/bash
mkdir -p gpurun_out
python tools/attn_ncu.py > gpurun_out/o_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_.*mma2 -s 2 -c 2 -o gpurun_out/o_attn python tools/attn_ncu.py > gpurun_out/o_ncu.log 2>&1
echo "ncu exit $?" > gpurun_out/o_status.log
cat gpurun_out/o_status.log gpurun_out/o_plain.log; tail -3 gpurun_out/o_ncu.log
