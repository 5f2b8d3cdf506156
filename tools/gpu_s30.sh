rm -f gpurun_out/s30_sweep.txt
for o in "" "--opt pull\ group=4" "--opt pull\ group=12" "--opt pull\ group=16" "--opt flush\ unroll=4" "--opt flush\ unroll=16" "--opt flush\ unroll=32" "--opt ring=metric" "--opt ring=metric --opt pull\ group=4" "--opt stage1=early" "--opt min\ blocks=4" "--opt debug\ skip=1" "--opt debug\ skip=2" "--opt debug\ skip=3"; do eval python bench.py --no-cpu-baseline --steps 20 $o 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['config'].get('plan_options'), 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4))
" >> gpurun_out/s30_sweep.txt; done; cat gpurun_out/s30_sweep.txt
