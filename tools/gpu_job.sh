#!/bin/bash
# The GPU job of the moment: `gpurun --timeout T -- 'bash tools/gpu_job.sh'`.  Overwritten between calls; results that matter are copied to profiles/.
mkdir -p gpurun_out
L=gpurun_out/r02_s7.log
: > $L
b() {
  python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']; print('$1', 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],4), 'M elem/s', round(d['value']/1e6,2), 'e2e M/s', round(d['e2e']['value']/1e6,2))
"
}
echo "== general path tests" >> $L
timeout 900 python -m pytest tests/test_gpu_general.py tests/test_gpu_fullsize.py -x -q -k "not thermal_128" 2>&1 | tail -3 >> $L
echo "== bench general (tensor | lanes)" >> $L
for w in ns leq2 le; do
  for j in tensor lanes; do
    timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-traffic --opt jacobian=$j 2>> gpurun_out/r02_s7.err | b "[$w $j]" >> $L
  done
done
timeout 400 python bench.py --workload maxwell --steps 10 --warmup 3 --no-cpu-baseline --no-traffic 2>> gpurun_out/r02_s7.err | b "[maxwell]" >> $L
echo "== launch lists" >> $L
for w in ns leq2; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 4 -c 2 --csv python bench.py --workload $w --traffic-child --steps 1 --warmup 3 2>/dev/null | grep -E "gen_|mrh_" | cut -c60-300 >> $L
done
echo "== thermal variants" >> $L
run() { v2=$(echo "$*" | sed 's/_/\\ /g'); eval timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-traffic $v2 2>> gpurun_out/r02_s7.err | b "[$*]" >> $L; }
run
run --opt prefetch=false
run --opt stage1=early
run --opt stage1=early --opt prefetch=false
timeout 300 python -m pytest tests/test_gpu_thermal.py -x -q 2>&1 | tail -2 >> $L
echo "== ncu" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gen_element -s 3 -c 1 -o gpurun_out/r02_s7_leq2_elem -f python bench.py --workload leq2 --n 40 --traffic-child --steps 1 --warmup 3 > /dev/null 2>> gpurun_out/r02_s7.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gen_element -s 3 -c 1 -o gpurun_out/r02_s7_ns_elem -f python bench.py --workload ns --n 64 --traffic-child --steps 1 --warmup 3 > /dev/null 2>> gpurun_out/r02_s7.err
cat $L
