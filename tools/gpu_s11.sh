python -m pytest tests/test_gpu_general.py -x -q 2>&1 | tail -3
for c in "ns 48" "ns 64" "le 64" "maxwell 48" "leq2 20" "thermal 64" "thq2 24"; do eval timeout 600 python tools/bench_general.py $c "batch\ elems=-1" >> gpurun_out/s11_gen_bench.jsonl 2>> gpurun_out/s11.err; done
cut -c1-200 gpurun_out/s11_gen_bench.jsonl; tail -3 gpurun_out/s11.err
ncu --set full --clock-control none --import-source on -k regex:gen_pull -s 2 -c 1 -o gpurun_out/s11_ns48_pull python tools/bench_general.py ns 48 steps=2 "batch elems=-1" > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gen_element -s 2 -c 1 -o gpurun_out/s11_leq2_elem python tools/bench_general.py leq2 20 steps=2 "batch elems=-1" > /dev/null 2>&1
