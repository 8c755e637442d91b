#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x -k "linear or tcgen05" > gpurun_out/am_pytest.log 2>&1; echo "pytest exit $?" > gpurun_out/am_status.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/am_bench.json 2> gpurun_out/am_bench.err; echo "bench exit $?" >> gpurun_out/am_status.log
cat gpurun_out/am_status.log; tail -n 2 gpurun_out/am_pytest.log | cut -c1-200; cut -c1-200 gpurun_out/am_bench.json
