#!/bin/bash
# experiment run 2: conv3to1 (merger layer6) thread-shape variants, contraction kernel vs cuBLAS / cuDNN on HEAD,
# the other BASELINE configuration shapes on one GPU (configs[3] per-GPU shard, the 1..24 view sweep at batch 128)
# (variant libraries first:  tools/build_variant.sh rpt2 svx_ops -DC31_RPT=2 ; rpt1 ; th8rpt4 = -DC31_RPT=4 -DC31_ROWS=8 ; th8rpt2 ; th8rpt1
#  -- the default was C31_ROWS=16 at the time.  Results: profiles/r2_conv3to1_variants.txt, r2_gemm_bench_v41.log, r2_config_sweep_v42.txt)
O=gpurun_out; mkdir -p $O
{
for v in "" _rpt2 _rpt1 _th8rpt4 _th8rpt2 _th8rpt1; do
  echo "variant '$v': $(SVX_ISOLATE=1 SVX_LIB_PATH=swinvox_b200/libswinvox_b200$v.so timeout 120 python tools/run_module.py merger 64 3 2>/dev/null | grep -E 'layer6|per call' | tr '\n' ' ')"
  SVX_LIB_PATH=swinvox_b200/libswinvox_b200$v.so timeout 120 python -m pytest tests/test_kernels.py -m gpu -q -p no:cacheprovider -k "single_output_fp32 or merger" 2>&1 | tail -1
done
} > $O/exp2_conv3to1.txt 2>&1
cat $O/exp2_conv3to1.txt
timeout 400 python tools/gemm_bench.py > $O/gemm_bench_r2.log 2>&1
tail -3 $O/gemm_bench_r2.log
{
echo "# python bench.py --batch B --views V --steps 5 --warmup 3 --no-eager --cpu-seconds 0 on one B200 (TF32 unless noted)"
for cfg in "32 20 tf32" "64 5 tf32" "128 1 tf32" "128 2 tf32" "128 3 tf32" "128 5 tf32" "128 8 tf32" "128 12 tf32" "128 16 tf32" "128 20 tf32" "128 24 tf32" "128 24 bf16" "1 1 tf32"; do
  set -- $cfg
  timeout 300 python bench.py --batch $1 --views $2 --dtype $3 --steps 5 --warmup 3 --no-eager --cpu-seconds 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
B,V=$1,$2
print(f'B={B} V={V} $3 ms/step {d[\"ms_per_step\"]:.2f} objects/s {d[\"value\"]:.1f} views/s {d[\"value\"]*V:.0f} e2e {d[\"e2e\"][\"value\"]:.1f} gemm frac {d[\"roofline\"][\"frac\"]:.3f} floor frac {d[\"roofline\"][\"step_frac_of_floor\"]:.3f} clocks {d[\"clocks\"][\"sm_mhz\"]} {d[\"clocks\"][\"reasons\"]}')
"
done
} > $O/config_sweep_r2.txt 2>&1
cat $O/config_sweep_r2.txt
