#!/bin/bash
# attention iteration: parity (short timeouts: a deadlocked kernel must not hold the box), timeline with kernel-internal marks
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_kernels.py tests/test_bench_shape_parity.py -x -q -k "attention or bs64" > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/d_pytest.log
VITK_LIB=dev timeout 120 python tools/step_timeline.py --detail --out gpurun_out/d_timeline > gpurun_out/d_timeline.log 2>&1; echo "timeline rc=$?"
head -12 gpurun_out/d_timeline.md
