rm -f gpurun_out/s31_sweep.txt
for o in "" "--opt stage1=early --opt stage2=early" "--opt stage1=early --opt stage2=early --opt pull\ group=4" "--opt stage1=early --opt stage2=early --opt flush\ unroll=32" "--opt stage1=early --opt stage2=early --opt debug\ skip=3" "--opt stage1=early --opt stage2=early --opt debug\ skip=1" "--opt stage1=early --opt stage2=early --opt ring=metric" "--opt stage1=early --opt stage2=early --opt min\ segment\ levels=20"; do eval python bench.py --no-cpu-baseline --steps 20 $o 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['config'].get('plan_options'), 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4))
" >> gpurun_out/s31_sweep.txt; done; cat gpurun_out/s31_sweep.txt
timeout 600 python -m pytest tests/test_gpu_thermal.py -x -q -k "metric_ring or regression or transient" 2>&1 | tail -2
