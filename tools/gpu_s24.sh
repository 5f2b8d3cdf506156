# columns long along the fastest id axis (16 x 8 instead of 8 x 16): shared-memory bank conflicts in the pull
rm -f gpurun_out/s24_sweep.txt
for o in "--opt min\ blocks=3 --opt pull\ group=8 --opt cta\ slots=444" "--opt min\ blocks=4 --opt pull\ group=8" "--opt ring=metric --opt min\ blocks=2" "--opt ring=full"; do eval python bench.py --no-cpu-baseline --steps 10 $o 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['config'].get('plan_options'), 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'chains', d['config']['chains'])
" >> gpurun_out/s24_sweep.txt; done; cat gpurun_out/s24_sweep.txt
