bash tools/gpu_s16.sh
bash tools/gpu_bounds_scan.sh 2>&1 | tee gpurun_out/s17_bounds_scan.txt
