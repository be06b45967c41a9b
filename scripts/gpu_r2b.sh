#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
echo "== host overhead"; timeout 300 python scripts/host_overhead.py 2>&1 | head -12
echo "== graph check"; timeout 300 python scripts/graph_check.py 2>&1 | tail -12
