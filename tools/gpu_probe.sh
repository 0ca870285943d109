#!/bin/bash
# timing probes of the window attention kernel (results are wrong with a probe set): where does the item time go?
for pr in ${PROBES:-0 1 2 3}; do
  SVX_WINATTN_PROBE=$pr timeout 300 python bench.py --no-eager --cpu-seconds 1 --steps 5 > gpurun_out/bench_probe$pr.json 2> gpurun_out/bench_probe$pr.err
  cp gpurun_out/op_breakdown.json gpurun_out/op_breakdown_probe$pr.json
done
python - <<PY
import json
for pr in [int(x) for x in '${PROBES:-0 1 2 3}'.split()]:
    try:
        ops = json.load(open(f"gpurun_out/op_breakdown_probe{pr}.json"))
        a = sorted([(o[0], round(o[1]*1e3)) for o in ops if o[0].endswith(".attn")])
        print("probe", pr, "attn ms", round(sum(x[1] for x in a)/1e3, 3), a[:2], a[2:4], a[4:5], a[-1:])
    except Exception as e:
        print(pr, "failed", e)
PY
