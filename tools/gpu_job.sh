#!/bin/bash
# The GPU job of the moment: `gpurun --gpus N --timeout T -- 'bash tools/gpu_job.sh N'`.  Overwritten between calls; results that matter are copied to profiles/.
# v18 kernel on N GPUs: thermal multirank parity kinds + weak bench line.
mkdir -p gpurun_out
L=gpurun_out/r02_s20.log
: > $L
N=${1:-2}
b() {
  python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']; print('$1', 'n_gpus', d['n_gpus'], d['scaling'], 'ms', round(d['ms_per_step'],4), 'median', round(d.get('ms_per_step_median',0),4), 'kernel_ms', round(r['kernel_ms'],4), 'G elem/s', round(d['value']/1e9,4), 'e2e M/s', round(d['e2e']['value']/1e6,1), d['config'].get('parallelism'))
"
}
echo "== multirank tests at world $N (thermal kinds)" >> $L
MRHYDE_B200_TEST_WORLD=$N timeout -k 5 300 python -m pytest tests/test_gpu_multirank.py -q -k thermal 2>&1 | tail -4 >> $L
echo "== bench" >> $L
timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-traffic 2>> gpurun_out/r02_s20.err | b "[thermal weak]" >> $L
cat $L
