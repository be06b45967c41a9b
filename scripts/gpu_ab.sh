#!/bin/bash
# A/B timing of tuning knobs: bash scripts/gpu_ab.sh "ENV=VAL" "ENV=VAL" ...   (each argument is one bench run; "-" = no env)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
summ='import json,sys; d=json.loads(sys.stdin.read()); print("step us %.2f bwd us %.2f fwd us %.2f" % (1e3*d["ms_per_step"], 1e3*d["roofline"]["avg_launch_ms"], 1e3*d["roofline_forward"]["avg_launch_ms"]))'
for round in 1 2; do
for cfg in "$@"; do
  echo "== $cfg"
  if [ "$cfg" == "-" ]; then envs=""; else envs="$cfg"; fi
  env $envs timeout 300 python bench.py --kernels-only --steps 300 --warmup 20 2>&1 | tail -1 | python -c "$summ"
done
done
