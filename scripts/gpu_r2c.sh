#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" == "1" ]; then
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-600
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 10 2>gpurun_out/bench_n1.err | tail -1 | tee gpurun_out/bench_n1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k,v in d.items(): print(k, json.dumps(v)[:400])
"
tail -5 gpurun_out/bench_n1.err
else
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline 2>gpurun_out/bench_n$N.err | tail -1 | tee gpurun_out/bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k,v in d.items(): print(k, json.dumps(v)[:400])
"
tail -8 gpurun_out/bench_n$N.err
echo "== pytest gpu (2-GPU exchange test)"; timeout 600 python -m pytest tests/test_peer_allreduce.py -m gpu -x -q 2>&1 | tail -4
fi
