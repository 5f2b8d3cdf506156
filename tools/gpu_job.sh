#!/bin/bash
# The GPU job of the moment: `gpurun --gpus N --timeout T -- 'bash tools/gpu_job.sh N'`.  Overwritten between calls; results that matter are copied to profiles/.
# Confirmation of the new default (L2 prefetch of the record streams): thermal + abi GPU tests, default bench line with traffic, transient and general-cell lines.
mkdir -p gpurun_out
O=gpurun_out
L=$O/r02_s15.log
: > $L
echo "== GPU tests: thermal, full size" >> $L
timeout -k 5 600 python -m pytest tests/test_gpu_thermal.py tests/test_gpu_fullsize.py -q -x 2>&1 | tail -3 >> $L
echo "== bench (default)" >> $L
timeout -k 5 400 python bench.py > $O/r02_v16_bench_thermal.json 2>> $O/r02_s15.err; cat $O/r02_v16_bench_thermal.json >> $L
echo "== bench, general cells" >> $L
timeout -k 5 300 python bench.py --perturb 0.1 --no-cpu-baseline --no-traffic >> $L 2>> $O/r02_s15.err
echo "== ncu full capture of the sweep kernel" >> $L
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:mrh_thermal_q1_3d -s 3 -c 1 -f -o $O/r02_v16_sweep python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-traffic > /dev/null 2>> $O/r02_s15.err
ls -la $O/r02_v16_sweep.ncu-rep >> $L 2>&1
tail -c 6000 $L
