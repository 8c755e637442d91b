#!/bin/bash
# what the driver runs at round end: the GPU tests, smoke(), the reference arm, the bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; echo "pytest exit $?" > gpurun_out/final_status.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/final_status.log
timeout 400 python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref exit $?" >> gpurun_out/final_status.log
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit $?" >> gpurun_out/final_status.log
cat gpurun_out/final_status.log; tail -n 3 gpurun_out/final_pytest.log | cut -c1-300; tail -n 2 gpurun_out/final_smoke.log | cut -c1-200
cut -c1-250 gpurun_out/final_ref.json; echo; cut -c1-250 gpurun_out/final_bench.json
# refreshed launch list of the training step (side stream off: serialised launches), only after everything above exited
VITK_KNOBS=8:1 timeout 200 python tools/step_profile.py > gpurun_out/final_plain.log 2>&1 && \
VITK_KNOBS=8:1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 320 -c 320 --csv --log-file gpurun_out/final_launches.csv python tools/step_profile.py > gpurun_out/final_ncu.log 2>&1
echo "ncu list exit $?" >> gpurun_out/final_status.log; tail -n 1 gpurun_out/final_status.log
