#!/bin/bash
# round 2, call B: LN forward (bulk-copy staged) + attention backward (ping-pong softmax groups): parity, bench, timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_bench_shape_parity.py -x -q -k "layernorm or attention or bs64" > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -4 gpurun_out/b_pytest.log
timeout 600 python bench.py --no-sustained --no-eager-baseline --no-cpu-baseline > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?"
VITK_LIB=dev timeout 300 python tools/step_timeline.py --out gpurun_out/b_timeline > gpurun_out/b_timeline.log 2>&1; echo "timeline rc=$?"
head -30 gpurun_out/b_timeline.md
