#!/bin/bash
# bash scripts/build_variant.sh NAME "-DFLAG ..."  -> gpurun_ab/NAME.so (another build of the same ABI, for R2L_ISP_LIB A/B timing)
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2
mkdir -p gpurun_ab/_obj_$name
pids=()
for src in raw2logit_b200/csrc/*.cu; do
  b=$(basename $src .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $flags -c -o gpurun_ab/_obj_$name/$b.o $src &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o gpurun_ab/$name.so gpurun_ab/_obj_$name/*.o
rm -rf gpurun_ab/_obj_$name
ls -la gpurun_ab/$name.so
