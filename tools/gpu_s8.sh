python -m pytest tests/test_gpu_general.py -x -q 2>&1 | tail -3
for c in "ns 48 batch\ elems=-1" "ns 64 batch\ elems=-1" "le 48 batch\ elems=-1" "le 64 batch\ elems=-1" "thermal 64 batch\ elems=-1" "leq2 20 batch\ elems=-1" "maxwell 48 batch\ elems=-1" "thq2 24 batch\ elems=-1"; do eval timeout 600 python tools/bench_general.py $c >> gpurun_out/s8_gen_bench.jsonl 2>> gpurun_out/s8.err; done
cut -c1-260 gpurun_out/s8_gen_bench.jsonl; tail -3 gpurun_out/s8.err
# headline kernel: sweep-plan parameter scan
for o in "" "--opt column\ elements=96" "--opt column\ elements=160" "--opt column\ elements=192" "--opt min\ chains=888" "--opt min\ chains=1184" "--opt min\ segment\ levels=4" "--opt min\ segment\ levels=16" "--opt cta\ slots=444"; do eval python bench.py --no-cpu-baseline --steps 10 $o 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['config'].get('plan_options'), 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'chains', d['config']['chains'], 'thr', d['config']['threads_per_block'], 'smem', d['config']['smem_bytes'])
" >> gpurun_out/s8_sweep.txt; done; cat gpurun_out/s8_sweep.txt
