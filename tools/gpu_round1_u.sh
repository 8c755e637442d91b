#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q --timeout 600 -x -k "not loss_curve" > gpurun_out/u_pytest.log 2>&1
echo "pytest exit $?" > gpurun_out/u_status.log
tail -4 gpurun_out/u_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err
echo "bench exit $?" >> gpurun_out/u_status.log
VITK_KNOBS=8:1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/u_bench_noside.json 2> gpurun_out/u_bench_noside.err
echo "bench noside exit $?" >> gpurun_out/u_status.log
cat gpurun_out/u_status.log; cut -c1-200 gpurun_out/u_bench.json; echo; cut -c1-200 gpurun_out/u_bench_noside.json; tail -15 gpurun_out/u_bench.err
