#!/bin/bash
# The GPU job of the moment: `gpurun --timeout T -- 'bash tools/gpu_job.sh'`.  Overwritten between calls; results that matter are copied to profiles/.
mkdir -p gpurun_out
L=gpurun_out/r02_s2.log
: > $L
b() {  # bench line -> short summary
  python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']; print('$1', 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],4), 'G elem/s', round(d['value']/1e9,3), 'chains', d['config'].get('chains'), 'smem', d['config'].get('smem_bytes'), 'e2e', round(d['e2e']['value']/1e6,1))
"
}
echo "== tests" >> $L
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_thermal.py -x -q 2>&1 | tail -6 >> $L
echo "== bench variants" >> $L
for v in "" "--opt flush=row" "--opt debug\ skip=1" "--opt debug\ skip=2" "--opt debug\ skip=3" "--opt flush=row --opt debug\ skip=1" "--opt ring=full" "--opt ring=full --opt flush=row" "--opt max\ blocks=2" "--opt pull\ group=12"; do
  eval timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $v 2>> gpurun_out/r02_s2.err | b "[$v]" >> $L
done
echo "== ncu" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mrh_thermal -s 3 -c 1 -o gpurun_out/r02_s2_thermal -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>> gpurun_out/r02_s2.err
ls -la gpurun_out/*.ncu-rep >> $L 2>&1
cat $L
