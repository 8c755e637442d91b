#!/bin/bash
# first GPU call: SIMT parity, tcgen05 probe, fp32 model parity
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.log 2>&1
python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "not tcgen05" --timeout 900 > gpurun_out/a_kernels_simt.log 2>&1
echo "kernels_simt exit $?" >> gpurun_out/a_status.log
timeout 600 python tools/tc_probe.py > gpurun_out/a_tc_probe.log 2>&1
echo "tc_probe exit $?" >> gpurun_out/a_status.log
timeout 1500 python -m pytest tests/test_gpu_model.py -m gpu -q -k "not bf16" --timeout 1200 > gpurun_out/a_model_fp32.log 2>&1
echo "model_fp32 exit $?" >> gpurun_out/a_status.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tcgen05" --timeout 300 > gpurun_out/a_kernels_tc.log 2>&1
echo "kernels_tc exit $?" >> gpurun_out/a_status.log
cat gpurun_out/a_status.log
tail -5 gpurun_out/a_kernels_simt.log gpurun_out/a_model_fp32.log
