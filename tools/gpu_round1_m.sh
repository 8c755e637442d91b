#!/bin/bash
mkdir -p gpurun_out
python tools/gemm_ncu.py > gpurun_out/m_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 4 -o gpurun_out/m_gemm python tools/gemm_ncu.py > gpurun_out/m_ncu.log 2>&1
echo "ncu exit $?" > gpurun_out/m_status.log
cat gpurun_out/m_status.log gpurun_out/m_plain.log; tail -5 gpurun_out/m_ncu.log
