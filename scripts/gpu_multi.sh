#!/bin/bash
# bash scripts/gpu_multi.sh N : bench + sweep + both train-step configs on N GPUs of one box (outputs under gpurun_out/)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
echo "== bench N=$N"; timeout 600 bash -c "$(declare -f run); N=$N; run 29511 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline" 2>gpurun_out/bench_n$N.err | tail -1 > gpurun_out/bench_n$N.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'e2e_fp32', d['e2e_fp32']['value'], 'eager', d['e2e_eager']['value'], d.get('exchange_check'))
PY
echo "== driver protocol (20 steps, 5 warmup)"; timeout 600 bash -c "$(declare -f run); N=$N; run 29512 bench.py --gpus $N --steps 20 --warmup 5 --kernels-only" 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'ms', d['ms_per_step'])"
echo "== sweep N=$N"; timeout 900 bash -c "$(declare -f run); N=$N; run 29513 scripts/sweep.py" 2>/dev/null | grep '^{' > gpurun_out/sweep_n$N.jsonl; wc -l gpurun_out/sweep_n$N.jsonl
echo "== train step microscopy N=$N"; timeout 600 bash -c "$(declare -f run); N=$N; run 29514 scripts/train_step_bench.py --task microscopy --batch 32" 2>/dev/null | grep '^{' | tee gpurun_out/train_micro_n$N.json
echo "== train step drone 256 N=$N"; timeout 600 bash -c "$(declare -f run); N=$N; run 29515 scripts/train_step_bench.py --task drone --batch 8 --size 256" 2>/dev/null | grep '^{' | tee gpurun_out/train_drone256_n$N.json
echo "== train step drone 2048 N=$N"; timeout 600 bash -c "$(declare -f run); N=$N; run 29516 scripts/train_step_bench.py --task drone --batch 1 --size 2048 --steps 10" 2>/dev/null | grep '^{' | tee gpurun_out/train_drone2048_n$N.json
