#!/bin/bash
# One GPU-box session: smoke + sanitizers + gpu tests + bench + ncu launch list + ncu full captures.
# Usage (from the build container): gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [quick]'
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
MODE="${1:-full}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.txt
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "== bench" ; timeout 600 python bench.py 2>&1 | tail -2 | tee gpurun_out/bench.json
if [ "$MODE" != "quick" ]; then
echo "== memcheck" ; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/memcheck.txt 2>&1 ; echo "memcheck exit $?" ; tail -4 gpurun_out/memcheck.txt
echo "== racecheck" ; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/racecheck.txt 2>&1 ; echo "racecheck exit $?" ; tail -6 gpurun_out/racecheck.txt
fi
echo "== ncu launches" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 12 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1 ; tail -2 gpurun_out/ncu_launch.log | cut -c1-300
echo "== ncu full" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:isp_ -s 12 -c 6 -o gpurun_out/prof -f python bench.py --steps 4 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1 ; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
