# class ring: parity tests, then launch-bound / pull-group / wave scan at 128^3
timeout 900 python -m pytest tests/test_gpu_thermal.py -x -q 2>&1 | tail -5 > gpurun_out/s22_tests.log; cat gpurun_out/s22_tests.log
rm -f gpurun_out/s22_sweep.txt
for o in "" "--opt min\ blocks=4 --opt pull\ group=8" "--opt min\ blocks=4 --opt pull\ group=4" "--opt min\ blocks=4 --opt pull\ group=8 --opt cta\ slots=592" "--opt min\ blocks=3 --opt pull\ group=8 --opt cta\ slots=444" "--opt min\ blocks=3 --opt pull\ group=28 --opt cta\ slots=444" "--opt min\ blocks=3 --opt pull\ group=8" "--opt min\ blocks=2" "--opt min\ blocks=2 --opt pull\ group=8" "--opt ring=metric --opt min\ blocks=2"; do eval python bench.py --no-cpu-baseline --steps 10 $o 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['config'].get('plan_options'), 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'chains', d['config']['chains'], 'thr', d['config']['threads_per_block'], 'smem', d['config']['smem_bytes'], 'halo', round(d['config']['elements_incl_halo']/d['config']['elements_per_gpu'],3))
" >> gpurun_out/s22_sweep.txt; done; cat gpurun_out/s22_sweep.txt
