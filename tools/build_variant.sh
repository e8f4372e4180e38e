#!/bin/bash
# usage: tools/build_variant.sh <name> <source.cu> [-D...]   -> variants/libcwa_<name>.so: the in-tree library with one source recompiled
# under other macros (tuning runs: CWA_LIB_PATH=variants/libcwa_<name>.so python bench.py ...).  *.so is git-ignored and travels with gpurun.
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
python -c "from coupledwateranimation_b200 import build as B; B.build()"
mkdir -p variants
base=$(basename "$src" .cu)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden "$@" -c -o variants/${base}_${name}.o coupledwateranimation_b200/csrc/$src
objs=$(ls coupledwateranimation_b200/build/*.o | grep -v "/${base}.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -o variants/libcwa_${name}.so $objs variants/${base}_${name}.o
echo variants/libcwa_${name}.so
