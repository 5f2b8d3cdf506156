#!/bin/bash
# The GPU job of the moment: `gpurun --gpus N --timeout T -- 'bash tools/gpu_job.sh N'`.  Overwritten between calls; results that matter are copied to profiles/.
# New coverage on the GPU: two-module blocks, point constraints, lumped scatter.
mkdir -p gpurun_out
L=gpurun_out/r02_s21.log
: > $L
echo "== new general-path cases + point constraints" >> $L
timeout -k 5 400 python -m pytest tests/test_gpu_general.py tests/test_gpu_thermal.py -q -k "thermoelastic or ns-thermal or lump or point_constraints" 2>&1 | tail -15 >> $L
echo "== smoke" >> $L
timeout -k 5 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 >> $L
cat $L
