#!/bin/bash
# The GPU job of the moment: `gpurun --gpus N --timeout T -- 'bash tools/gpu_job.sh N'`.  Overwritten between calls; results that matter are copied to profiles/.
# Final single-GPU pass of round 2: whole GPU suite, smoke, both bench arms, the other workloads, launch list, one full ncu capture.
mkdir -p gpurun_out
O=gpurun_out
L=$O/r02_final.log
: > $L
echo "== pytest -m gpu" >> $L
timeout -k 5 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 >> $L
echo "== smoke" >> $L
timeout -k 5 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 >> $L
echo "== bench (default)" >> $L
timeout -k 5 400 python bench.py > $O/r02_final_bench_thermal.json 2>> $O/r02_final.err; cat $O/r02_final_bench_thermal.json >> $L
echo "== bench --impl reference" >> $L
timeout -k 5 400 python bench.py --impl reference > $O/r02_final_bench_reference.json 2>> $O/r02_final.err; cat $O/r02_final_bench_reference.json >> $L
for w in le leq2 ns maxwell; do
  echo "== bench --workload $w" >> $L
  timeout -k 5 300 python bench.py --workload $w --no-cpu-baseline > $O/r02_final_bench_$w.json 2>> $O/r02_final.err; cat $O/r02_final_bench_$w.json >> $L
done
echo "== bench thermal, general cells (--perturb 0.1)" >> $L
timeout -k 5 300 python bench.py --perturb 0.1 --no-cpu-baseline > $O/r02_final_bench_thermal_general_cells.json 2>> $O/r02_final.err; cat $O/r02_final_bench_thermal_general_cells.json >> $L
echo "== ncu launch list" >> $L
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-traffic > /dev/null 2>> $O/r02_final.err
grep -c . $O/r02_final_launches.csv >> $L
echo "== ncu full capture of the sweep kernel" >> $L
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:mrh_thermal_q1_3d -s 3 -c 1 -f -o $O/r02_final_sweep python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-traffic > /dev/null 2>> $O/r02_final.err
ls -la $O/r02_final_sweep.ncu-rep >> $L 2>&1
tail -c 3000 $L
