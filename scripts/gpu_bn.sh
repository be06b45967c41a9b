#!/bin/bash
cd "$(dirname "$0")/.."
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for e in "-" "R2L_ISP_BN_SPLIT=1"; do
  echo "== bench $e"; if [ "$e" == "-" ]; then ee=""; else ee="$e"; fi
  env $ee timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('step', d['ms_per_step'], 'bn_tail', d['bn_tail'], 'e2e', d['e2e']['value'])"
done
echo "== train step micro (host-bound small batch)"; timeout 300 python scripts/train_step_bench.py --task microscopy --batch 32 2>/dev/null | grep '^{' | cut -c1-260
