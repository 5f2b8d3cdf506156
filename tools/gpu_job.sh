#!/bin/bash
# The GPU job of the moment: `gpurun --gpus N --timeout T -- 'bash tools/gpu_job.sh N'`.  Overwritten between calls; results that matter are copied to profiles/.
# CTA-shared sub-expression values: the last warp evaluates the sweep-axis sub-expressions of the next step during the pull.
mkdir -p gpurun_out
O=gpurun_out
L=$O/r02_s19.log
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
run() { timeout -k 5 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-traffic "${@:2}" 2>> $O/r02_s19.err | b "$1" >> $L; }
run "[shared z values + column cache (default)]"
run "[column cache in registers only (v17)]" --opt "column cache=registers"
run "[default again]"
echo "== bench (default, full line)" >> $L
timeout -k 5 400 python bench.py > $O/r02_v18_bench_thermal.json 2>> $O/r02_s19.err; cat $O/r02_v18_bench_thermal.json >> $L
cat $L
