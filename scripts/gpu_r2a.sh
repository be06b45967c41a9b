#!/bin/bash
# round-2 first look: gen5 vs gen6 timing, ncu full capture of each backward (one launch)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
summ='import json,sys; d=json.loads(sys.stdin.read()); print("step ms %.4f bwd ms %.4f fwd ms %.4f e2e %.0f u16 %.0f" % (d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline_forward"]["avg_launch_ms"], d["e2e"]["value"], d["e2e_uint16"]["value"]))'
echo "== bench gen6"; timeout 300 python bench.py --no-cpu-baseline --steps 100 2>&1 | tail -1 | tee gpurun_out/bench_gen6.json | python -c "$summ"
echo "== bench gen5"; R2L_ISP_BWD_GEN=5 timeout 300 python bench.py --no-cpu-baseline --steps 100 2>&1 | tail -1 | tee gpurun_out/bench_gen5.json | python -c "$summ"
echo "== ncu gen6"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:isp_backward -s 6 -c 1 -o gpurun_out/prof_gen6 -f python bench.py --steps 4 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_gen6.log 2>&1; tail -2 gpurun_out/ncu_gen6.log
echo "== ncu gen5 + fwd"; R2L_ISP_BWD_GEN=5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:isp_ -s 12 -c 2 -o gpurun_out/prof_gen5 -f python bench.py --steps 4 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_gen5.log 2>&1; tail -2 gpurun_out/ncu_gen5.log
ls -la gpurun_out
