#!/bin/bash
# A/B of bench.py argument sets: bash scripts/gpu_ab2.sh "--batch 32 --sets 1" "--batch 32 --sets 8" ...
cd "$(dirname "$0")/.."
summ='import json,sys; d=json.loads(sys.stdin.read()); print("step ms %.4f bwd ms %.4f fwd ms %.4f" % (d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline_forward"]["avg_launch_ms"]))'
for cfg in "$@"; do
  echo "== $cfg"
  timeout 300 python bench.py --no-cpu-baseline --steps 200 $cfg 2>&1 | tail -1 | python -c "$summ"
done
