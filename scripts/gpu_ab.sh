#!/bin/bash
# A/B timing of tuning knobs: bash scripts/gpu_ab.sh "ENV=VAL" "ENV=VAL" ...   (each argument is one bench run)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
summ='import json,sys; d=json.loads(sys.stdin.read()); print("step ms %.4f bwd ms %.4f fwd ms %.4f" % (d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline_forward"]["avg_launch_ms"]))'
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --no-cpu-baseline --steps 200 2>&1 | tail -1 | python -c "$summ"
done
