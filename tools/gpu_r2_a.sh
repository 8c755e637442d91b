#!/bin/bash
# round 2, call A: full GPU test suite, bench, device-side timeline, compute-sanitizer smoke
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 600 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
VITK_LIB=dev timeout 300 python tools/step_timeline.py --out gpurun_out/a_timeline > gpurun_out/a_timeline.log 2>&1; echo "timeline rc=$?"
for tool in memcheck synccheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --kernel-name-exclude regex:at:: --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/a_san_$tool.log 2>&1
  echo "sanitizer $tool rc=$?"
done
tail -3 gpurun_out/a_pytest.log
