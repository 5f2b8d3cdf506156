# where does the step time go: builds without global stores / without element work (timing experiments only)
rm -f gpurun_out/s23_sweep.txt
B="--opt min\ blocks=3 --opt pull\ group=8 --opt cta\ slots=444"
for o in "$B" "$B --opt debug\ skip=1" "$B --opt debug\ skip=2" "$B --opt debug\ skip=3" "--opt ring=metric --opt min\ blocks=2 --opt debug\ skip=1" "--opt ring=metric --opt min\ blocks=2 --opt debug\ skip=2" "--opt ring=metric --opt min\ blocks=2 --opt debug\ skip=3"; do eval python bench.py --no-cpu-baseline --steps 10 $o 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['config'].get('plan_options'), 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'chains', d['config']['chains'])
" >> gpurun_out/s23_sweep.txt; done; cat gpurun_out/s23_sweep.txt
