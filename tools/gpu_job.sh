#!/bin/bash
# The GPU job of the moment: `gpurun --gpus N --timeout T -- 'bash tools/gpu_job.sh N'`.  Overwritten between calls; results that matter are copied to profiles/.
mkdir -p gpurun_out
L=gpurun_out/r02_s11.log
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
echo "== multirank tests (thermal kinds)" >> $L
timeout -k 5 400 python -m pytest tests/test_gpu_multirank.py -q -k "thermal" 2>&1 | tail -6 >> $L
echo "== bench" >> $L
tr() { timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline "${@:2}" 2>> gpurun_out/r02_s11.err; }
tr 29511 | b "[thermal weak p2p + in-kernel push]" >> $L
tr 29512 --opt "halo push=false" | b "[thermal weak p2p, push in the halo kernel]" >> $L
tr 29513 --scaling strong | b "[thermal strong p2p + in-kernel push]" >> $L
cat $L
