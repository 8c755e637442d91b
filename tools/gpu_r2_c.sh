#!/bin/bash
# round 2, call C: patch-embed TMA kernels, LN forward (bulk-copy staged), attention backward (ping-pong): parity, bench, timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_patch_embed.py -q > gpurun_out/c_pe.log 2>&1; echo "patch-embed pytest rc=$?"; tail -15 gpurun_out/c_pe.log
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_bench_shape_parity.py tests/test_eval_metrics.py -x -q -k "layernorm or attention or bs64 or uint8" > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/c_pytest.log
timeout 600 python bench.py --no-sustained --no-eager-baseline --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"
VITK_LIB=dev timeout 300 python tools/step_timeline.py --detail --out gpurun_out/c_timeline > gpurun_out/c_timeline.log 2>&1; echo "timeline rc=$?"
head -30 gpurun_out/c_timeline.md
