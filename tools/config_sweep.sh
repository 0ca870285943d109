for cfg in "64 5" "32 20" "128 1" "16 24" "128 3" "1 1"; do
  set -- $cfg
  timeout 300 python bench.py --batch $1 --views $2 --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
B,V=$1,$2
print(f'B={B} V={V} ms/step {d[\"ms_per_step\"]:.2f} objects/s {d[\"value\"]:.1f} views/s {d[\"value\"]*V:.0f} e2e {d[\"e2e\"][\"value\"]:.1f} gemm frac {d[\"roofline\"][\"frac\"]:.3f} floor frac {d[\"roofline\"][\"step_frac_of_floor\"]:.3f}')
"
done
