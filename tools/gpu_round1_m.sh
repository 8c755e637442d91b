#!/bin/bash
mkdir -p gpurun_out
python tools/gemm_ncu.py > gpurun_out/m_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 4 -o gpurun_out/m_gemm3 python tools/gemm_ncu.py > gpurun_out/m_ncu.log 2>&1
echo "ncu gemm exit $?" > gpurun_out/m_status.log
python tools/attn_ncu.py > gpurun_out/o_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_.*tc_kernel -s 2 -c 2 -o gpurun_out/o_attn_tc2 python tools/attn_ncu.py > gpurun_out/o_ncu.log 2>&1
echo "ncu attn exit $?" >> gpurun_out/m_status.log
VITK_KNOBS=8:1 python tools/step_profile.py > gpurun_out/t_plain.log 2>&1 && \
VITK_KNOBS=8:1 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 330 --csv --log-file gpurun_out/t_launches2.csv python tools/step_profile.py > gpurun_out/t_ncu.log 2>&1
echo "ncu list exit $?" >> gpurun_out/m_status.log
cat gpurun_out/m_status.log
