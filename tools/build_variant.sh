#!/bin/bash
# usage: tools/build_variant.sh <name> "<extra nvcc flags>"  -> dispersal-simulation_b200/ab/<name>.so (working tree + flags, for tools/ab.sh)
set -e
name=$1; extra=$2
root=$(cd "$(dirname "$0")/.." && pwd)/dispersal-simulation_b200
mkdir -p "$root/ab"
cd "$root"
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fopenmp -ccbin /usr/bin/g++ $extra -shared -o ab/$name.so csrc/*.cu -lcudart -lnccl -lgomp 2>&1 | grep -E "error" || true
echo "built $name.so with [$extra]"
