#!/bin/bash
# The GPU job of the moment: `gpurun --gpus N --timeout T -- 'bash tools/gpu_job.sh N'`.  Overwritten between calls; results that matter are copied to profiles/.
mkdir -p gpurun_out
L=gpurun_out/r02_s12.log
: > $L
N=${1:-4}
b() {
  python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']; print('$1', 'n_gpus', d['n_gpus'], d['scaling'], 'ms', round(d['ms_per_step'],4), 'median', round(d.get('ms_per_step_median',0),4), 'kernel_ms', round(r['kernel_ms'],4), 'G elem/s', round(d['value']/1e9,4), 'e2e M/s', round(d['e2e']['value']/1e6,1), d['config'].get('parallelism'))
"
}
echo "== multirank tests at world $N (thermal, thermal_nopush, maxwell)" >> $L
MRHYDE_B200_TEST_WORLD=$N timeout -k 5 300 python -m pytest tests/test_gpu_multirank.py -q -k "thermal and not nccl and not overlap or maxwell" 2>&1 | tail -6 >> $L
echo "== bench" >> $L
tr() { n=$1; timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline --no-traffic "${@:3}" 2>> gpurun_out/r02_s12.err | tee -a gpurun_out/r02_s12_scale.jsonl; }
tr $N 29511 | b "[thermal weak]" >> $L
tr 2 29512 | b "[thermal weak]" >> $L
tr $N 29513 --scaling strong | b "[thermal strong]" >> $L
tr $N 29514 --opt "halo push=false" | b "[thermal weak, push in the halo kernel]" >> $L
cat $L
