rm -f gpurun_out/s38_sweep.txt
for o in "" "--opt tables=literal" "" "--opt tables=literal"; do eval python bench.py --no-cpu-baseline --steps 20 $o 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['config'].get('plan_options'), 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4))
" >> gpurun_out/s38_sweep.txt; done; cat gpurun_out/s38_sweep.txt
