python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload ns --no-cpu-baseline > gpurun_out/s15_bench_ns.json 2> gpurun_out/s15_ns.err; tail -2 gpurun_out/s15_ns.err
python bench.py --workload le --no-cpu-baseline > gpurun_out/s15_bench_le.json 2> gpurun_out/s15_le.err; tail -2 gpurun_out/s15_le.err
cut -c1-200 gpurun_out/s15_bench_ns.json gpurun_out/s15_bench_le.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s15_ns96_launches.csv python bench.py --workload ns --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
