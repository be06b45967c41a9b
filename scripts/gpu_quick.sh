#!/bin/bash
# Quick GPU iteration: gpu tests + one bench line + one full ncu capture of the fused kernels.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
echo "== bench" ; timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step ms', d['ms_per_step'], 'bwd ms', d['roofline']['avg_launch_ms'], 'fwd ms', d['roofline_forward']['avg_launch_ms'], 'e2e', d['e2e']['value'], 'u16', d['e2e_uint16']['value'], 'frac', d['roofline_step_frac'])"
if [ "${1:-}" != "noprof" ]; then
echo "== ncu full" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:isp_ -s 12 -c 2 -o gpurun_out/prof -f python bench.py --steps 4 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1 ; tail -1 gpurun_out/ncu_full.log
fi
