#!/bin/bash
# The GPU job of the moment: `gpurun --gpus N --timeout T -- 'bash tools/gpu_job.sh N'`.  Overwritten between calls; results that matter are copied to profiles/.
# mrhyde_b200_set_initial on the GPU against the oracle (the last seconds of the round's GPU budget).
mkdir -p gpurun_out
timeout -k 2 18 python -m pytest tests/test_gpu_general.py -q -x -k set_initial_as_a_whole -p no:cacheprovider 2>&1 | tail -3 > gpurun_out/r02_s24.log
cat gpurun_out/r02_s24.log
