# start-up stagger of the first wave (desynchronise the phases of co-resident CTAs)
rm -f gpurun_out/s28_sweep.txt
for o in "" "--opt stagger\ ns=2000" "--opt stagger\ ns=4000" "--opt stagger\ ns=6000" "--opt stagger\ ns=9000" "--opt stagger\ ns=15000" "--opt stagger\ ns=6000 --opt pull\ group=28"; do eval python bench.py --no-cpu-baseline --steps 10 $o 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['config'].get('plan_options'), 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4))
" >> gpurun_out/s28_sweep.txt; done; cat gpurun_out/s28_sweep.txt
