python -m pytest tests/test_gpu_general.py -x -q -k "full_size or weighted_mass" 2>&1 | tail -6
for c in "leq2 32" "leq2 48"; do eval timeout 900 python tools/bench_general.py $c steps=5 >> gpurun_out/s16_gen_bench.jsonl 2>> gpurun_out/s16.err; done
cut -c1-230 gpurun_out/s16_gen_bench.jsonl; tail -3 gpurun_out/s16.err
