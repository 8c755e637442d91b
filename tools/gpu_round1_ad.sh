#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x -k "wgrad or streamk or attention" > gpurun_out/ad_pytest.log 2>&1; echo "pytest exit $?" > gpurun_out/ad_status.log
timeout 300 python tools/attn_bench.py > gpurun_out/ad_attn.log 2>&1; echo "attn exit $?" >> gpurun_out/ad_status.log
rm -f gpurun_out/ad_cublas.log
for k in "" "15:1"; do
  ONLY=wgrad VITK_KNOBS=$k timeout 200 python tools/cublas_yardstick.py >> gpurun_out/ad_cublas.log 2>&1; echo "yardstick [$k] exit $?" >> gpurun_out/ad_status.log
done
timeout 300 python tools/knob_ab.py 12:0 12:1 --rounds 4 --steps 10 > gpurun_out/ad_knob.log 2>&1; echo "knob exit $?" >> gpurun_out/ad_status.log
cat gpurun_out/ad_status.log; tail -n 3 gpurun_out/ad_pytest.log | cut -c1-300; grep "variant 0" gpurun_out/ad_attn.log; grep -v "^shape" gpurun_out/ad_cublas.log | cut -c1-110; cat gpurun_out/ad_knob.log
