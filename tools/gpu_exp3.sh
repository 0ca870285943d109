#!/bin/bash
# validation run of the activation-buffer reuse + 64-byte raw rows: GPU tier, bench line, the large ends of the view sweep
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -8 > $O/exp3_tests.txt
cat $O/exp3_tests.txt
timeout 200 python bench.py --steps 10 --no-eager --cpu-seconds 0 2> $O/exp3_bench.err | cut -c1-700
python tools/op_sum.py reuse
{
for cfg in "128 12 tf32" "128 16 tf32" "128 20 tf32" "128 24 tf32" "128 24 bf16"; do
  set -- $cfg
  timeout 300 python bench.py --batch $1 --views $2 --dtype $3 --steps 5 --warmup 3 --no-eager --cpu-seconds 0 2> $O/exp3_sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
B,V=$1,$2
print(f'B={B} V={V} $3 ms/step {d[\"ms_per_step\"]:.2f} objects/s {d[\"value\"]:.1f} views/s {d[\"value\"]*V:.0f} e2e {d[\"e2e\"][\"value\"]:.1f} gemm frac {d[\"roofline\"][\"frac\"]:.3f} floor frac {d[\"roofline\"][\"step_frac_of_floor\"]:.3f} clocks {d[\"clocks\"][\"sm_mhz\"]} {d[\"clocks\"][\"reasons\"]}')
" || tail -2 $O/exp3_sweep.err | cut -c1-300
  nvidia-smi --query-gpu=memory.used --format=csv,noheader
done
} > $O/exp3_sweep.txt 2>&1
cat $O/exp3_sweep.txt
