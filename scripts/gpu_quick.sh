#!/bin/bash
# Quick GPU iteration: smoke, gpu tests, one bench line (+ the previous backward generation for comparison),
# one full ncu capture of the fused kernels.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== smoke" ; timeout 180 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
summ='import json,sys; d=json.loads(sys.stdin.read()); print("step ms", d["ms_per_step"], "bwd ms", d["roofline"]["avg_launch_ms"], "fwd ms", d["roofline_forward"]["avg_launch_ms"], "e2e", d["e2e"]["value"], "u16", d["e2e_uint16"]["value"], "frac", d["roofline_step_frac"])'
echo "== bench" ; timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench.json | python -c "$summ"
if [ "${2:-}" == "ab" ]; then
echo "== bench gen4" ; R2L_ISP_BWD_GEN=4 timeout 600 python bench.py --no-cpu-baseline --steps 100 2>&1 | tail -1 | tee gpurun_out/bench_gen4.json | python -c "$summ"
fi
if [ "${3:-}" == "launches" ]; then
echo "== ncu launch list" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1 ; grep -c isp_ gpurun_out/launches.csv
fi
if [ "${1:-}" != "noprof" ]; then
echo "== ncu full" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:isp_ -s 12 -c 2 -o gpurun_out/prof -f python bench.py --steps 4 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1 ; tail -1 gpurun_out/ncu_full.log
fi
