#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x -k "wgrad or streamk or attention" > gpurun_out/ac_pytest.log 2>&1; echo "pytest exit $?" > gpurun_out/ac_status.log
timeout 300 python tools/attn_bench.py > gpurun_out/ac_attn.log 2>&1; echo "attn exit $?" >> gpurun_out/ac_status.log
rm -f gpurun_out/ac_cublas.log
for k in "" "14:1" "13:1" "13:1,14:1"; do
  ONLY=wgrad VITK_KNOBS=$k timeout 200 python tools/cublas_yardstick.py >> gpurun_out/ac_cublas.log 2>&1; echo "yardstick [$k] exit $?" >> gpurun_out/ac_status.log
done
timeout 300 python tools/knob_ab.py 13:0,14:0 13:1,14:1 --rounds 4 --steps 10 > gpurun_out/ac_knob.log 2>&1; echo "knob exit $?" >> gpurun_out/ac_status.log
cat gpurun_out/ac_status.log; tail -n 3 gpurun_out/ac_pytest.log | cut -c1-300; grep "variant 0" gpurun_out/ac_attn.log; grep -v "^shape" gpurun_out/ac_cublas.log | cut -c1-110; cat gpurun_out/ac_knob.log
