#!/bin/bash
# round-2 profile capture of the default bench command: launch list, full capture of one forward + one backward launch
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r02_bench.json; cut -c1-300 gpurun_out/r02_bench.json
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --warmup 3 --kernels-only > gpurun_out/ncu_launch.log 2>&1; grep -c isp_ gpurun_out/r02_launches.csv
echo "== full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:isp_ -s 12 -c 2 -o gpurun_out/r02_prof -f python bench.py --steps 4 --warmup 4 --kernels-only > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
echo "== bn full"; timeout 900 ncu --set full --clock-control none -k regex:"isp_|bn_" -s 20 -c 6 -o gpurun_out/r02_prof_bn -f python scripts/bn_step_time.py > gpurun_out/ncu_bn.log 2>&1; tail -1 gpurun_out/ncu_bn.log
