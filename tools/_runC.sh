for mode in "" "SVX_SLAB_KDN=1"; do
  echo "== spin $mode"
  env SVX_LIB_PATH=$PWD/tools/probes/libswinvox_b200_spin.bin SVX_ISOLATE=1 $mode python tools/run_module.py merger 64 3 3 2>&1 | grep -E "merger"
  echo "== spin+profile $mode"
  env SVX_LIB_PATH=$PWD/tools/probes/libswinvox_b200_spinprof.bin SVX_ISOLATE=1 $mode python tools/run_module.py merger 64 3 1 2>&1 | grep -E "slab profile" | sed -n 2,3p
done
echo "== spin bench"
SVX_LIB_PATH=$PWD/tools/probes/libswinvox_b200_spin.bin python bench.py 2>/dev/null | cut -c1-220
echo "== default bench"
python bench.py 2>/dev/null | cut -c1-220
