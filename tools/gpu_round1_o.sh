#!/bin/bash
mkdir -p gpurun_out
python tools/attn_ncu.py > gpurun_out/o_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_.*tc_kernel -s 2 -c 2 -o gpurun_out/o_attn_tc python tools/attn_ncu.py > gpurun_out/o_ncu.log 2>&1
echo "ncu exit $?" > gpurun_out/o_status.log
cat gpurun_out/o_status.log gpurun_out/o_plain.log; tail -n 3 gpurun_out/o_ncu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err
cut -c1-200 gpurun_out/o_bench.json; grep -o '"extra": {[^}]*}' gpurun_out/o_bench.json
