#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_checkpoint_interop.py tests/test_grad_scaler.py -m gpu -q --timeout 300 -x -k "not loss_curve" > gpurun_out/ai_pytest.log 2>&1; echo "pytest exit $?" > gpurun_out/ai_status.log
timeout 300 python tools/knob_ab.py noprezero:0 noprezero:1 --rounds 5 --steps 10 > gpurun_out/ai_knob.log 2>&1; echo "knob exit $?" >> gpurun_out/ai_status.log
cat gpurun_out/ai_status.log; tail -n 5 gpurun_out/ai_pytest.log | cut -c1-300; cat gpurun_out/ai_knob.log
