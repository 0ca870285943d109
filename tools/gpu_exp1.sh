#!/bin/bash
# experiment run: wait back-off variants of the attention kernel, merger layer5 64-byte rows, refiner FC tiles
# (variant libraries first, in the build container:  tools/build_variant.sh bo40all svx_winattn -DSVX_WU_BACKOFF_NS=40 ; bo40slack / bo100slack:
#  -DSVX_WU_BACKOFF_NS=40|100 -DSVX_WU_BACKOFF_TAGS=0x22u.  Result: profiles/README.md, "What the round-2 captures showed")
O=gpurun_out; mkdir -p $O
run() {  # label, lib
  SVX_LIB_PATH=$2 timeout 200 python bench.py --steps 5 --no-eager --cpu-seconds 0 2> $O/exp1_$1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"
  python tools/op_sum.py $1
}
{
run base swinvox_b200/libswinvox_b200.so
run bo40all swinvox_b200/libswinvox_b200_bo40all.so
run bo40slack swinvox_b200/libswinvox_b200_bo40slack.so
run bo100slack swinvox_b200/libswinvox_b200_bo100slack.so
} > $O/exp1.txt 2>&1
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider -k "merger or pipeline or refiner or modules" 2>&1 | tail -5 >> $O/exp1.txt
cat $O/exp1.txt
