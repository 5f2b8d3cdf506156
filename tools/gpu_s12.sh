python -m pytest tests/test_gpu_general.py -x -q 2>&1 | tail -3
for c in "leq2 20" "leq2 24"; do eval timeout 600 python tools/bench_general.py $c "batch\ elems=-1" >> gpurun_out/s12_gen_bench.jsonl 2>> gpurun_out/s12.err; done
cut -c1-200 gpurun_out/s12_gen_bench.jsonl; tail -3 gpurun_out/s12.err
ncu --set full --clock-control none --import-source on -k regex:gen_element -s 2 -c 1 -o gpurun_out/s12_leq2_elem python tools/bench_general.py leq2 20 steps=2 "batch elems=-1" > /dev/null 2>&1
