# ncu full capture of the metric-ring sweep kernel at 128^3 (default launch bounds and min blocks = 2)
ncu --set full --clock-control none --import-source on -k regex:mrh_thermal -s 3 -c 1 -o gpurun_out/s21_metric_mb4 python bench.py --no-cpu-baseline --steps 2 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:mrh_thermal -s 3 -c 1 -o gpurun_out/s21_metric_mb2 python bench.py --no-cpu-baseline --steps 2 --warmup 3 --opt "min blocks=2" > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
