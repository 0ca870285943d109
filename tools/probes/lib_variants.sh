#!/bin/bash
# builds variants of the whole library with extra -D switches: tools/probes/lib_variants.sh "<flags 0>" "<flags 1>" ...
# -> tools/probes/build/libsvx_var_<i>.so, selected at run time with SVX_LIB_PATH (same sources, never another backend)
set -e
cd "$(dirname "$0")/../.."
SRC="swinvox_b200/csrc/svx_gemm.cu swinvox_b200/csrc/svx_mlp.cu swinvox_b200/csrc/svx_ops.cu swinvox_b200/csrc/svx_io.cu swinvox_b200/csrc/svx_api.cu"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared"
mkdir -p tools/probes/build
i=0
for v in "$@"; do
  nvcc $FLAGS $v $SRC -o tools/probes/build/libsvx_var_$i.so 2>/dev/null &
  i=$((i+1))
done
wait
ls -la tools/probes/build
