# 8-GPU weak-scaling line of the headline workload (bounded)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/s35_thermal_8.json 2> gpurun_out/s35.err; echo rc=$?; cut -c1-230 gpurun_out/s35_thermal_8.json; tail -2 gpurun_out/s35.err
