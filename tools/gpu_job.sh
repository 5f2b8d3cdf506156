#!/bin/bash
# The GPU job of the moment: `gpurun --gpus N --timeout T -- 'bash tools/gpu_job.sh N'`.  Overwritten between calls; results that matter are copied to profiles/.
# Column cache: one-coordinate sub-expressions of the source kept per thread along an extruded chain.
mkdir -p gpurun_out
O=gpurun_out
L=$O/r02_s17.log
: > $L
b() {
  python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']; print('$1', 'ms', round(d['ms_per_step'],4), 'median', round(d.get('ms_per_step_median',0),4), 'kernel_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],4), 'G elem/s', round(d['value']/1e9,4))
"
}
echo "== GPU tests: thermal, full size" >> $L
timeout -k 5 600 python -m pytest tests/test_gpu_thermal.py tests/test_gpu_fullsize.py -q -x 2>&1 | tail -3 >> $L
run() { timeout -k 5 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-traffic "${@:2}" 2>> $O/r02_s17.err | b "$1" >> $L; }
run "[column cache (default)]"
run "[column cache, no gather prefetch]" --opt "prefetch=false"
run "[column cache, early stage1, no gather prefetch]" --opt "stage1=early" --opt "prefetch=false"
run "[column cache (default) again]"
echo "== bench (default, full line)" >> $L
timeout -k 5 400 python bench.py > $O/r02_v17_bench_thermal.json 2>> $O/r02_s17.err; cat $O/r02_v17_bench_thermal.json >> $L
echo "== ncu" >> $L
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:mrh_thermal_q1_3d -s 3 -c 1 -f -o $O/r02_s17_sweep python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-traffic > /dev/null 2>> $O/r02_s17.err
ls -la $O/r02_s17_sweep.ncu-rep >> $L 2>&1
cat $L
