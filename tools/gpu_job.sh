#!/bin/bash
# The GPU job of the moment: `gpurun --gpus N --timeout T -- 'bash tools/gpu_job.sh N'`.  Overwritten between calls; results that matter are copied to profiles/.
# Last GPU run of the round: the whole GPU suite on the final code.
mkdir -p gpurun_out
L=gpurun_out/r02_final3.log
: > $L
timeout -k 5 215 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 >> $L
cat $L
