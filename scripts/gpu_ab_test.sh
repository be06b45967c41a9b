#!/bin/bash
# In-call A/B of the in-tree build against gpurun_ab/base.so (two alternating rounds), then the GPU test suite on the in-tree build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_ab.sh "R2L_ISP_LIB=gpurun_ab/base.so" "-" "$@" 2>&1 | tee gpurun_out/ab.txt
echo "== pytest gpu" ; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/ab_pytest.txt
