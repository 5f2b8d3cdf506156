#!/bin/bash
# The GPU job of the moment: `gpurun --gpus N --timeout T -- 'bash tools/gpu_job.sh N'`.  Overwritten between calls; results that matter are copied to profiles/.
# New coverage on the GPU: previous-step / previous-stage Jacobians, fix zero rows; transient sweep-kernel tests (the seeds touched them).
mkdir -p gpurun_out
L=gpurun_out/r02_s22.log
: > $L
timeout -k 5 400 python -m pytest tests/test_gpu_general.py tests/test_gpu_thermal.py -q -k "previous_step or fix or transient or dirk or bwe or bdf or adjoint" 2>&1 | tail -12 >> $L
cat $L
