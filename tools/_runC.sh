for lib in nomma nomma_epi; do
echo "== $lib"; env SVX_LIB_PATH=$PWD/tools/probes/libswinvox_b200_$lib.bin SVX_ISOLATE=1 timeout 120 python tools/run_module.py merger 64 3 3 2>&1 | grep "merger.layer" | sed -n '1p;5p;7p'
done
