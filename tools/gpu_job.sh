#!/bin/bash
# The GPU job of the moment: `gpurun --gpus N --timeout T -- 'bash tools/gpu_job.sh N'`.  Overwritten between calls; results that matter are copied to profiles/.
# Final single-GPU pass on the v17 kernel: whole GPU suite, smoke, default bench line, launch list (general-path lines, reference arm and the full ncu capture: r02_final_* and r02_v17_*).
mkdir -p gpurun_out
O=gpurun_out
L=$O/r02_final2.log
: > $L
echo "== pytest -m gpu" >> $L
timeout -k 5 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 >> $L
echo "== smoke" >> $L
timeout -k 5 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 >> $L
echo "== bench (default)" >> $L
timeout -k 5 400 python bench.py > $O/r02_final2_bench_thermal.json 2>> $O/r02_final2.err; cat $O/r02_final2_bench_thermal.json >> $L
echo "== ncu launch list" >> $L
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_final2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-traffic > /dev/null 2>> $O/r02_final2.err
grep -c . $O/r02_final2_launches.csv >> $L
tail -c 3000 $L
