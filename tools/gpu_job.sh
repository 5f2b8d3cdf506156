#!/bin/bash
# The GPU job of the moment: `gpurun --gpus N --timeout T -- 'bash tools/gpu_job.sh N'`.  Overwritten between calls; results that matter are copied to profiles/.
mkdir -p gpurun_out
L=gpurun_out/r02_s9.log
: > $L
N=${1:-8}
b() {
  python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']; print('$1', 'n_gpus', d['n_gpus'], d['scaling'], 'ms', round(d['ms_per_step'],4), 'median', round(d.get('ms_per_step_median',0),4), 'kernel_ms', round(r['kernel_ms'],4), 'G elem/s', round(d['value']/1e9,4), 'e2e M/s', round(d['e2e']['value']/1e6,1), d['config'].get('parallelism'))
    open('gpurun_out/r02_s9_lines.jsonl','a').write(l)
"
}
tr() { n=$1; timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline "${@:3}" 2>> gpurun_out/r02_s9.err; }
tr 8 29511 | b "[thermal weak]" >> $L
tr 8 29512 --scaling strong | b "[thermal strong]" >> $L
tr 4 29513 | b "[thermal weak]" >> $L
tr 8 29514 --workload maxwell --scaling strong | b "[maxwell strong]" >> $L
tr 8 29515 --workload leq2 --scaling strong | b "[leq2 strong]" >> $L
MRHYDE_B200_HALO_TRANSPORT=nccl tr 8 29516 | b "[thermal weak nccl]" >> $L
cat $L
