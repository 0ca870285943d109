#!/bin/bash
# role timers of the merger's slab kernel (instrumented build of the same sources); run on the GPU box
set -x
for mode in "" "SVX_SLAB_KDN=1"; do
  echo "== $mode"
  env SVX_LIB_PATH=$PWD/tools/probes/libswinvox_b200_prof.bin SVX_ISOLATE=1 $mode python tools/run_module.py merger 64 3 1 2>&1 | grep -E "slab profile|merger.layer" | tail -16
done
