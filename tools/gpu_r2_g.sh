#!/bin/bash
# round 2, call G: full GPU suite, three-arm same-process A/B of the split tail (default / depth threshold 12 / off), full bench,
# launch list of one whole step (cudaProfilerStart/Stop around step 3), DRAM traffic of its GEMM launches, --set full of the
# first backward block's GEMMs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g3_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/g3_pytest.log
timeout 200 python tools/knob_ab.py 9:0 9:1 --rounds 4 --steps 10 > gpurun_out/g3_ab.log 2>&1; echo "ab rc=$?"; tail -4 gpurun_out/g3_ab.log
timeout 600 python bench.py > gpurun_out/g3_bench.json 2> gpurun_out/g3_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/g3_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'clocks',d['clocks'])
print('roofline',d['roofline']['achieved'],d['roofline']['frac'],d['roofline']['gemm_share_of_step'])
x=d['extra']
print({k:x[k] for k in x if k.startswith('bs') or k.startswith('frozen') or k.startswith('mfu') or k.startswith('speedup') or k.startswith('eager')})
print(x.get('sustained')); print(x.get('hbm_kernels')); print(d.get('cpu_baseline'))
PY
tail -14 gpurun_out/g3_bench.err
CMD="python tools/step_profile.py"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/g3_launches.csv $CMD > gpurun_out/g3_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:gemm_tc --csv --log-file gpurun_out/g3_gemm_traffic.csv $CMD > gpurun_out/g3_ncu_traffic.log 2>&1; echo "gemm traffic rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc -s 44 -c 12 -f -o gpurun_out/g3_gemm_fb $CMD > gpurun_out/g3_ncu_gemm_fb.log 2>&1; echo "gemm fwd+bwd rc=$?"
ls -la gpurun_out/g3_*
