#!/bin/bash
mkdir -p gpurun_out
for k in "" "7:1" "7:2" "7:3"; do
  echo "=== knobs $k" >> gpurun_out/q_gemm.log
  VITK_KNOBS="$k" timeout 300 python tools/gemm_bench.py >> gpurun_out/q_gemm.log 2>&1
done
cat gpurun_out/q_gemm.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
VITK_NO_PDL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/q_bench_nopdl.json 2> gpurun_out/q_bench_nopdl.err
cut -c1-200 gpurun_out/q_bench.json; echo; cut -c1-200 gpurun_out/q_bench_nopdl.json; tail -14 gpurun_out/q_bench.err
