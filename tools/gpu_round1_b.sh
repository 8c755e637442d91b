#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 1200 > gpurun_out/b_pytest.log 2>&1
echo "pytest exit $?" > gpurun_out/b_status.log
python __graft_entry__.py --smoke > gpurun_out/b_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/b_status.log
python bench.py --steps 10 --warmup 3 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
echo "bench exit $?" >> gpurun_out/b_status.log
cat gpurun_out/b_status.log; tail -3 gpurun_out/b_pytest.log; cat gpurun_out/b_smoke.log | tail -4; cat gpurun_out/b_bench.json; tail -30 gpurun_out/b_bench.err
