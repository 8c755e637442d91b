#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ah_cublas.log
timeout 200 python tools/cublas_yardstick.py >> gpurun_out/ah_cublas.log 2>&1
for k in "13:1" "13:70" "14:1"; do
  ONLY=wgrad VITK_KNOBS=$k timeout 200 python tools/cublas_yardstick.py >> gpurun_out/ah_cublas.log 2>&1
done
timeout 300 python tools/knob_ab.py 13:0 13:1 13:70 --rounds 4 --steps 10 > gpurun_out/ah_knob.log 2>&1
grep -v "^shape" gpurun_out/ah_cublas.log | cut -c1-150; cat gpurun_out/ah_knob.log
