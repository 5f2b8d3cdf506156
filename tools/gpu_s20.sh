# metric ring: parity tests, then A/B timing against the full ring and a first parameter scan
timeout 900 python -m pytest tests/test_gpu_thermal.py -x -q 2>&1 | tail -15 > gpurun_out/s20_tests.log; cat gpurun_out/s20_tests.log
rm -f gpurun_out/s20_sweep.txt
for o in "" "--opt ring=full" "--opt min\ blocks=3" "--opt min\ blocks=2" "--opt cta\ slots=592" "--opt cta\ slots=592 --opt min\ chains=1184" "--opt column\ elements=64 --opt threads=128" "--opt column\ elements=64 --opt threads=128 --opt cta\ slots=888" "--opt pull\ patterns=0"; do eval python bench.py --no-cpu-baseline --steps 10 $o 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['config'].get('plan_options'), 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'chains', d['config']['chains'], 'thr', d['config']['threads_per_block'], 'smem', d['config']['smem_bytes'], 'halo', round(d['config']['elements_incl_halo']/d['config']['elements_per_gpu'],3))
" >> gpurun_out/s20_sweep.txt; done; cat gpurun_out/s20_sweep.txt
