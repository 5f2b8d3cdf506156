#!/bin/bash
# The GPU job of the moment: `gpurun --timeout T -- 'bash tools/gpu_job.sh'`.  Overwritten between calls; results that matter are copied to profiles/.
mkdir -p gpurun_out
L=gpurun_out/r02_s6.log
: > $L
b() {
  python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']; print('$1', 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],4), 'M elem/s', round(d['value']/1e6,2), 'e2e M/s', round(d['e2e']['value']/1e6,2), 'traffic', r.get('traffic'), 'alg', r.get('algorithmic_bytes_per_launch'))
"
}
echo "== general path tests (tensor-core contraction)" >> $L
timeout 900 python -m pytest tests/test_gpu_general.py tests/test_gpu_fullsize.py -x -q -k "not thermal_128" 2>&1 | tail -5 >> $L
echo "== bench" >> $L
for w in ns leq2 le maxwell; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-traffic 2>> gpurun_out/r02_s6.err | b "[$w]" >> $L
done
echo "== launch list ns" >> $L
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 4 -c 4 --csv python bench.py --workload ns --traffic-child --steps 1 --warmup 3 2>/dev/null | grep -E "gen_|mrh_" | cut -c1-300 >> $L
echo "== launch list leq2" >> $L
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 4 -c 4 --csv python bench.py --workload leq2 --traffic-child --steps 1 --warmup 3 2>/dev/null | grep -E "gen_|mrh_" | cut -c1-300 >> $L
echo "== ncu leq2 element kernel" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gen_element -s 3 -c 1 -o gpurun_out/r02_s6_leq2_elem -f python bench.py --workload leq2 --n 40 --traffic-child --steps 1 --warmup 3 > /dev/null 2>> gpurun_out/r02_s6.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gen_element -s 3 -c 1 -o gpurun_out/r02_s6_ns_elem -f python bench.py --workload ns --n 64 --traffic-child --steps 1 --warmup 3 > /dev/null 2>> gpurun_out/r02_s6.err
ls -la gpurun_out/*.ncu-rep >> $L 2>&1
cat $L
