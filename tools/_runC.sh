for parts in 1 5; do for mode in "" "SVX_SLAB_NO_PAIR=1"; do
echo "== parts=$parts $mode"; env SVX_LIB_PATH=$PWD/tools/probes/libswinvox_b200_parts$parts.bin $mode SVX_ISOLATE=1 timeout 120 python tools/run_module.py merger 64 3 3 2>&1 | grep "merger.layer" | sed -n '1p;5p'
done; done
SVX_LIB_PATH=$PWD/tools/probes/libswinvox_b200_parts5.bin timeout 180 python -m pytest tests/test_kernels.py -m gpu -q -x -k "slab" 2>&1 | tail -2
