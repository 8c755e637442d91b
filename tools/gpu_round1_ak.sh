#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_full_size_properties.py -m gpu -q --timeout 600 -x > gpurun_out/ak_pytest.log 2>&1; echo "pytest exit $?" > gpurun_out/ak_status.log
timeout 300 python tools/knob_ab.py 12:0 12:1 --rounds 5 --steps 10 > gpurun_out/ak_knob.log 2>&1; echo "knob exit $?" >> gpurun_out/ak_status.log
ONLY=dgrad timeout 200 python tools/cublas_yardstick.py > gpurun_out/ak_cublas.log 2>&1
cat gpurun_out/ak_status.log; tail -n 5 gpurun_out/ak_pytest.log | cut -c1-300; cat gpurun_out/ak_knob.log; grep -v "^shape" gpurun_out/ak_cublas.log | cut -c1-110
