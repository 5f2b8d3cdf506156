# halo sum cost on 2 GPUs vs NCCL point-to-point channel count (bounded runs)
for e in "" "NCCL_MIN_P2P_NCHANNELS=8" "NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32" "NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32"; do
  env $e timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/s36.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$e', 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'value', round(d['value']/1e9,2))
" >> gpurun_out/s36_halo.txt
done; cat gpurun_out/s36_halo.txt
