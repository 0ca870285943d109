#!/bin/bash
# early stage release of the attention kernel (-DSVX_WU_EARLY_RELEASE=1 variant): timing, kernel tests, stress, pipeline parity
# (variant library first:  tools/build_variant.sh early svx_winattn -DSVX_WU_EARLY_RELEASE=1 -DSVX_WU_LOAD_WARPS=2.  Result: no change)
O=gpurun_out; mkdir -p $O
E=swinvox_b200/libswinvox_b200_early.so
{
timeout 60 python tools/winattn_time.py 192 2>&1 | tail -1
SVX_LIB_PATH=$E timeout 60 python tools/winattn_time.py 192 2>&1 | tail -1
(SVX_LIB_PATH=$E timeout 100 python -m pytest tests/test_kernels.py -m gpu -q -p no:cacheprovider -k "window_attention" 2>&1 | tail -2) &
(SVX_LIB_PATH=$E timeout 100 python -m pytest tests/test_modules.py -m gpu -q -p no:cacheprovider -k "default or stages_13" 2>&1 | tail -2) &
(SVX_LIB_PATH=$E timeout 100 python tools/stress_winattn.py 300 64 3 tf32 2>&1 | grep -E "^stress|Error|FAILED|stall") &
wait
} > $O/exp4.txt 2>&1
cat $O/exp4.txt
