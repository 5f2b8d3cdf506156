# round-1 final measurements of the headline path: full GPU suite, bench line, launch list, full ncu capture, smoke
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/s32_gpu_tests.log; cat gpurun_out/s32_gpu_tests.log
python bench.py > gpurun_out/s32_bench_thermal.json 2> gpurun_out/s32_bench.err; tail -c 1500 gpurun_out/s32_bench_thermal.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s32_bench_reference.json 2>> gpurun_out/s32_bench.err; tail -c 600 gpurun_out/s32_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s32_launches.csv python bench.py --no-cpu-baseline --steps 2 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:mrh_thermal -s 3 -c 1 -o gpurun_out/s32_class python bench.py --no-cpu-baseline --steps 2 --warmup 3 > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
