# 2-GPU weak-scaling line of the headline workload (bounded: a hang must not burn the budget)
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/s34_thermal_2.json 2> gpurun_out/s34.err; echo rc=$?; cut -c1-230 gpurun_out/s34_thermal_2.json; tail -2 gpurun_out/s34.err
