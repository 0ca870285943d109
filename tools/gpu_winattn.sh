#!/bin/bash
# window attention bring-up: tests under a hang guard, then A/B per-op times (new tcgen05 kernel vs the mma.sync one)
set -x
tag=${1:-wa}
timeout 300 python -m pytest tests/test_kernels.py tests/test_kernels_bf16.py -m gpu -k "window" -q 2>&1 | tail -30 > gpurun_out/wa_tests_$tag.log
cat gpurun_out/wa_tests_$tag.log | tail -12
if grep -q "failed\|error" gpurun_out/wa_tests_$tag.log; then exit 0; fi
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/gpu_tests_$tag.log; tail -3 gpurun_out/gpu_tests_$tag.log
timeout 300 python bench.py --no-eager --cpu-seconds 2 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cp gpurun_out/op_breakdown.json gpurun_out/op_breakdown_$tag.json
SVX_WINATTN_MMASYNC=1 timeout 300 python bench.py --no-eager --cpu-seconds 2 > gpurun_out/bench_${tag}_legacy.json 2> gpurun_out/bench_${tag}_legacy.err; cp gpurun_out/op_breakdown.json gpurun_out/op_breakdown_${tag}_legacy.json
timeout 300 python bench.py --dtype bf16 --views 5 --no-eager --cpu-seconds 2 > gpurun_out/bench_${tag}_bf16v5.json 2> gpurun_out/bench_${tag}_bf16v5.err; cp gpurun_out/op_breakdown.json gpurun_out/op_breakdown_${tag}_bf16v5.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"winattn_umma" -c 3 \
    -o gpurun_out/prof_winattn_$tag python tools/run_module.py encoder 64 3 1 > gpurun_out/ncu_winattn_$tag.log 2>&1
python - <<PY
import json
for t in ("$tag", "${tag}_legacy", "${tag}_bf16v5"):
    try:
        d = json.load(open(f"gpurun_out/bench_{t}.json")); ops = json.load(open(f"gpurun_out/op_breakdown_{t}.json"))
        print(t, round(d["value"]), round(d["ms_per_step"], 2), "attn ms", round(sum(o[1] for o in ops if o[0].endswith(".attn")), 3))
    except Exception as e:
        print(t, "failed", e)
PY
