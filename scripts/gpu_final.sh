#!/bin/bash
# Final verification of a build on one GPU box: smoke, the GPU test suite, both bench arms under the driver's protocol.
# Usage: gpurun --timeout 900 -- 'bash scripts/gpu_final.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/final_smoke.txt
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/final_pytest_gpu.txt
echo "== bench reference arm" ; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/final_bench_ref.json | cut -c1-400
echo "== bench" ; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/final_bench.json | cut -c1-1200
