for o in "--opt threads=192" "--opt threads=224" "--opt threads=256" "--opt column\ elements=112" "--opt column\ elements=144" "--opt column\ elements=64 --opt threads=96" "--opt column\ elements=64 --opt threads=128" "--opt sweep\ axis=0" "--opt sweep\ axis=1"; do eval python bench.py --no-cpu-baseline --steps 10 $o 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['config'].get('plan_options'), 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'chains', d['config']['chains'], 'thr', d['config']['threads_per_block'], 'smem', d['config']['smem_bytes'], 'halo', round(d['config']['elements_incl_halo']/d['config']['elements_per_gpu'],3))
" >> gpurun_out/s19_sweep.txt; done; cat gpurun_out/s19_sweep.txt
