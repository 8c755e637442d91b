#!/bin/bash
# round 2 profiles: launch list of one training step, DRAM traffic of every GEMM launch, --set full of the top kernels
mkdir -p gpurun_out
CMD="python tools/step_profile.py"
timeout 120 $CMD > gpurun_out/n_plain.log 2>&1; echo "plain rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 330 --csv --log-file gpurun_out/n_launches.csv $CMD > gpurun_out/n_ncu_launches.log 2>&1; echo "launch list rc=$?"
head -3 gpurun_out/n_launches.csv
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tc -s 292 -c 146 --csv --log-file gpurun_out/n_gemm_traffic.csv $CMD > gpurun_out/n_ncu_traffic.log 2>&1; echo "gemm traffic rc=$?"
cap() {  # name, regex, skip, count
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/n_$1 $CMD > gpurun_out/n_ncu_$1.log 2>&1; echo "$1 rc=$?"
}
cap attn attn_ 24 2
cap attn_bwd attn_bwd 12 2
cap ln ln_fwd 50 2
cap gemm gemm_tc 147 8
cap gemm_bwd gemm_tc 231 10
cap patch patch_embed 2 2
ls -la gpurun_out/n_*.ncu-rep 2>/dev/null
