#!/bin/bash
mkdir -p gpurun_out
python tools/attn_bench.py > gpurun_out/i_attn.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err
echo "bench exit $?" > gpurun_out/i_status.log
cat gpurun_out/i_status.log gpurun_out/i_attn.log; cut -c1-600 gpurun_out/i_bench.json
