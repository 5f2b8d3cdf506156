python -m pytest tests/test_gpu_general.py -x -q 2>&1 | tail -3
for c in "ns 48" "ns 64" "le 64" "maxwell 48" "leq2 20"; do eval timeout 600 python tools/bench_general.py $c "batch\ elems=-1" >> gpurun_out/s9_gen_bench.jsonl 2>> gpurun_out/s9.err; done
cut -c1-200 gpurun_out/s9_gen_bench.jsonl
ncu --set full --clock-control none --import-source on -k regex:gen_element -s 2 -c 1 -o gpurun_out/s9_ns48_elem python tools/bench_general.py ns 48 steps=2 "batch elems=-1" > /dev/null 2>&1
