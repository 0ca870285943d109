#!/bin/bash
# An instrumented / experimental build of ONE translation unit next to the product library (same sources, other -D flags):
#   tools/build_variant.sh <name> <file.cu> -DFOO=1 ...   ->  swinvox_b200/libswinvox_b200_<name>.so
# Run it with SVX_LIB_PATH=swinvox_b200/libswinvox_b200_<name>.so (never the default; the product loads libswinvox_b200.so).
set -e
name=$1; src=$2; shift 2
R=$(cd "$(dirname "$0")/.." && pwd)
O=$R/swinvox_b200/csrc/_obj
[ -f $O/svx_api.o ] || python $R/__graft_entry__.py build
base=$(basename $src .cu)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c $R/swinvox_b200/csrc/$base.cu -o $O/${base}_$name.o
objs=""
for f in svx_gemm svx_mlp svx_ops svx_winattn svx_io svx_api; do
  if [ $f = $base ]; then objs="$objs $O/${base}_$name.o"; else objs="$objs $O/$f.o"; fi
done
nvcc -shared -gencode arch=compute_100a,code=sm_100a $objs -o $R/swinvox_b200/libswinvox_b200_$name.so
echo built $R/swinvox_b200/libswinvox_b200_$name.so
