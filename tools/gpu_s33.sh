# 2-GPU: multirank parity tests + weak-scaling bench lines (thermal headline, ns) after the halo-sum change
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q 2>&1 | tail -3 > gpurun_out/s33_tests.log; cat gpurun_out/s33_tests.log
python bench.py --no-cpu-baseline > gpurun_out/s33_thermal_1.json 2> gpurun_out/s33.err; cut -c1-220 gpurun_out/s33_thermal_1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/s33_thermal_2.json 2>> gpurun_out/s33.err; cut -c1-220 gpurun_out/s33_thermal_2.json
tail -3 gpurun_out/s33.err
