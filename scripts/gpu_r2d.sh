#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== model test"; timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -15
echo "== ref on gpu"; timeout 600 python scripts/ref_on_gpu.py 2>&1 | grep -v Warning | tee gpurun_out/ref_on_gpu.jsonl
echo "== sweep"; timeout 900 python scripts/sweep.py 2>&1 | tee gpurun_out/sweep_n1.jsonl | cut -c1-220
