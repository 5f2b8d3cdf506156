( time python bench.py --workload leq2 --n 64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s18_bench_leq2_64.json 2> gpurun_out/s18_64.err ) 2>&1 | tail -3
tail -2 gpurun_out/s18_64.err; cut -c1-300 gpurun_out/s18_bench_leq2_64.json
python bench.py --workload leq2 > gpurun_out/s18_bench_leq2_48.json 2> gpurun_out/s18_48.err; tail -2 gpurun_out/s18_48.err; cut -c1-300 gpurun_out/s18_bench_leq2_48.json
