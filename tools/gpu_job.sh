#!/bin/bash
# The GPU job of the moment: `gpurun --gpus N --timeout T -- 'bash tools/gpu_job.sh N'`.  Overwritten between calls; results that matter are copied to profiles/.
# Sweep kernel: next step's connectivity requested before the element work (stage1=early) + L2 prefetch of its gathers before the barrier.
mkdir -p gpurun_out
O=gpurun_out
L=$O/r02_s13.log
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
echo "== thermal GPU tests (variants)" >> $L
timeout -k 5 300 python -m pytest tests/test_gpu_thermal.py -q -x -k "variant or early or option" 2>&1 | tail -3 >> $L
run() { timeout -k 5 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-traffic "${@:2}" 2>> $O/r02_s13.err | b "$1" >> $L; }
run "[late stage1 (default), unconditional step-record load]"
run "[stage1=early + prefetch before the barrier]" --opt "stage1=early"
run "[stage1=early, no prefetch]" --opt "stage1=early" --opt "prefetch=false"
run "[late again]"
run "[early again]" --opt "stage1=early"
echo "== ncu of the early variant" >> $L
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:mrh_thermal_q1_3d -s 3 -c 1 -f -o $O/r02_s13_sweep_early python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-traffic --opt "stage1=early" > /dev/null 2>> $O/r02_s13.err
ls -la $O/r02_s13_sweep_early.ncu-rep >> $L 2>&1
cat $L
