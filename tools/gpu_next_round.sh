# First GPU runs of the next round (nothing here has run yet).  Usage: gpurun --gpus 2 --timeout 600 -- 'bash tools/gpu_next_round.sh 2'
# 1. parity of the overlapped halo exchange (option "overlap halo"), bounded so a hang costs little
# 2. weak-scaling bench line with and without it
N=${1:-2}
# 0. paths that were only checked on the host so far: transient builds of the metric ring
MRHYDE_B200_TEST_UNVALIDATED=1 timeout 300 python -m pytest tests/test_gpu_thermal.py -x -q -k metric_ring_transient 2>&1 | tail -4
MRHYDE_B200_TEST_OVERLAP=1 timeout 240 python -m pytest tests/test_gpu_multirank.py -x -q -k thermal 2>&1 | tail -4
for e in 0 1; do
  MRHYDE_B200_OVERLAP_HALO=$e timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$e \
    bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/next_overlap$e.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('overlap=$e', 'n_gpus', d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'G elem/s', round(d['value']/1e9,2))
"
done
