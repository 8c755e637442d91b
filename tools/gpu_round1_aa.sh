#!/bin/bash
# validation of HEAD + refreshed evidence: GPU tests (loss curves excluded: final call runs them), smoke, default bench,
# torch-eager baseline on the same box, launch list, ncu --set full of the HBM-bound kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -m gpu -q --timeout 600 -x -k "not loss_curve" > gpurun_out/aa_pytest.log 2>&1
echo "pytest exit $?" > gpurun_out/aa_status.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/aa_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/aa_status.log
timeout 400 python bench.py > gpurun_out/aa_bench.json 2> gpurun_out/aa_bench.err
echo "bench exit $?" >> gpurun_out/aa_status.log
timeout 300 python tools/torch_eager_baseline.py > gpurun_out/aa_eager.json 2> gpurun_out/aa_eager.err
echo "eager exit $?" >> gpurun_out/aa_status.log
VITK_KNOBS=8:1 timeout 200 python tools/step_profile.py > gpurun_out/aa_plain.log 2>&1 && \
VITK_KNOBS=8:1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 330 --csv --log-file gpurun_out/aa_launches.csv python tools/step_profile.py > gpurun_out/aa_ncu.log 2>&1
echo "ncu list exit $?" >> gpurun_out/aa_status.log
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:ln_(fwd|bwd)_kernel" -s 72 -c 6 -f -o gpurun_out/aa_ln python tools/step_profile.py > gpurun_out/aa_ncu_ln.log 2>&1
echo "ncu ln exit $?" >> gpurun_out/aa_status.log
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:adam_kernel|sumsq_partial_kernel|focal_kernel" -s 3 -c 3 -f -o gpurun_out/aa_opt python tools/step_profile.py > gpurun_out/aa_ncu_opt.log 2>&1
echo "ncu opt exit $?" >> gpurun_out/aa_status.log
cat gpurun_out/aa_status.log; tail -n 3 gpurun_out/aa_pytest.log | cut -c1-300; tail -n 3 gpurun_out/aa_smoke.log | cut -c1-300
cut -c1-220 gpurun_out/aa_bench.json; echo; cat gpurun_out/aa_eager.json; tail -n 3 gpurun_out/aa_eager.err
