# 2-GPU: overlapped halo exchange -- parity test, then the weak-scaling bench line (bounded)
timeout 300 python -m pytest tests/test_gpu_multirank.py -x -q -k thermal 2>&1 | tail -4
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s39_thermal_2.json 2> gpurun_out/s39.err; echo rc=$?; cut -c1-220 gpurun_out/s39_thermal_2.json; tail -2 gpurun_out/s39.err
