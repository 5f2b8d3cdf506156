# defaults after the class ring + row flush, then plan-parameter scan
rm -f gpurun_out/s26_sweep.txt
for o in "" "--opt stage1=late" "--opt min\ segment\ levels=40" "--opt min\ segment\ levels=20" "--opt min\ chains=1700" "--opt column\ elements=64" "--opt column\ elements=64 --opt max\ blocks=5" "--opt column\ elements=256 --opt threads=256" "--opt pull\ group=12" "--opt pull\ group=28" "--opt threads=192" "--opt ring=metric" "--opt ring=metric --opt pull\ group=8"; do eval python bench.py --no-cpu-baseline --steps 10 $o 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['config'].get('plan_options'), 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'chains', d['config']['chains'], 'seg', d['config']['segments'], 'thr', d['config']['threads_per_block'], 'smem', d['config']['smem_bytes'], 'halo', round(d['config']['elements_incl_halo']/d['config']['elements_per_gpu'],3))
" >> gpurun_out/s26_sweep.txt; done; cat gpurun_out/s26_sweep.txt
ncu --set full --clock-control none --import-source on -k regex:mrh_thermal -s 3 -c 1 -o gpurun_out/s26_class python bench.py --no-cpu-baseline --steps 2 --warmup 3 > /dev/null 2>&1
