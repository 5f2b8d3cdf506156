#!/bin/bash
# The GPU job of the moment: `gpurun --gpus N --timeout T -- 'bash tools/gpu_job.sh N'`.  Overwritten between calls; results that matter are copied to profiles/.
# C++ host example on the GPU (examples/host_assemble.cpp through the C ABI, host buffers) against the oracle.
mkdir -p gpurun_out
timeout -k 3 40 python -m pytest tests/test_gpu_thermal.py -q -x -k cpp_host_example 2>&1 | tail -4 > gpurun_out/r02_s23.log
cat gpurun_out/r02_s23.log
