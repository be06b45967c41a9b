#!/bin/bash
# Final evidence of a build on one GPU box: smoke + GPU tests + both bench arms (gpu_final.sh), the ncu launch list of the
# bench command and one full capture (forward + backward launch) with raw and source pages.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_final.sh
echo "== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 20 --warmup 3 --kernels-only > gpurun_out/ncu_launch.log 2>&1; grep -c isp_ gpurun_out/final_launches.csv
echo "== full"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:isp_ -s 12 -c 2 -o gpurun_out/final_prof -f python bench.py --steps 4 --warmup 4 --kernels-only > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
ncu -i gpurun_out/final_prof.ncu-rep --page raw --csv > gpurun_out/final_ncu_raw.csv 2>/dev/null
ncu -i gpurun_out/final_prof.ncu-rep --page source --csv --print-source sass --kernel-name regex:isp_backward5 > gpurun_out/final_bwd_source.csv 2>/dev/null
ncu -i gpurun_out/final_prof.ncu-rep --page source --csv --print-source sass --kernel-name regex:isp_forward3 > gpurun_out/final_fwd_source.csv 2>/dev/null
