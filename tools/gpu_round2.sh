#!/bin/bash
# Round-2 evidence run on the GPU box:  gpurun --timeout 2400 -- 'bash tools/gpu_round2.sh r2_vNN [quick|full]'
set -x
tag=${1:-r2_vXX}
mode=${2:-full}
rm -f gpurun_out/parity_$tag.txt
SVX_PARITY_LOG=gpurun_out/parity_$tag.txt timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/gpu_tests_$tag.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
cp gpurun_out/op_breakdown.json gpurun_out/op_breakdown_$tag.json
timeout 600 python bench.py --dtype bf16 --views 5 --no-eager --cpu-seconds 3 > gpurun_out/bench_${tag}_bf16v5.json 2> gpurun_out/bench_${tag}_bf16v5.err
cp gpurun_out/op_breakdown.json gpurun_out/op_breakdown_${tag}_bf16v5.json
timeout 600 python bench.py --views 5 --no-eager --cpu-seconds 3 > gpurun_out/bench_${tag}_tf32v5.json 2> gpurun_out/bench_${tag}_tf32v5.err
if [ "$mode" = "full" ]; then
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2> gpurun_out/bench_${tag}_reference.err
  # the stress run of the attention kernel: graph replays with two other streams loading the copy engines and the SMs
  for cfg in "17 1 tf32" "33 1 tf32" "64 3 tf32" "17 1 bf16" "64 5 bf16"; do
    timeout 300 python tools/stress_winattn.py 300 $cfg 2>&1 | grep -E "^stress|Error|FAILED|stall" >> gpurun_out/stress_$tag.txt
  done
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4400 --csv \
      --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-eager --cpu-seconds 0 > gpurun_out/ncu_launch_$tag.log 2>&1
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4400 --csv \
      --log-file gpurun_out/launches_${tag}_bf16v5.csv python bench.py --dtype bf16 --views 5 --steps 2 --warmup 3 --no-eager --cpu-seconds 0 > gpurun_out/ncu_launch_${tag}_bf16v5.log 2>&1
  # one full capture of the kernels the step time sits in (first launches of the second warm-up step)
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:"winattn_umma_kernel|mlp_fused_kernel|conv3_slab_kernel|gemm_tf32_kernel" --launch-skip 400 -c 60 \
      -o gpurun_out/ncu_full_$tag python bench.py --steps 1 --warmup 3 --no-eager --cpu-seconds 0 > gpurun_out/ncu_full_$tag.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:"winattn_umma_kernel|gemm_bf16_kernel|lnrows_bf16_kernel" --launch-skip 200 -c 40 \
      -o gpurun_out/ncu_full_${tag}_bf16v5 python bench.py --dtype bf16 --views 5 --steps 1 --warmup 3 --no-eager --cpu-seconds 0 > gpurun_out/ncu_full_${tag}_bf16v5.log 2>&1
  gzip -f gpurun_out/launches_$tag.csv gpurun_out/launches_${tag}_bf16v5.csv
  ls -la gpurun_out/*.ncu-rep
fi
cat gpurun_out/stress_$tag.txt
cut -c1-300 gpurun_out/bench_${tag}_bf16v5.json; tail -3 gpurun_out/bench_${tag}_bf16v5.err; cut -c1-300 gpurun_out/bench_${tag}_tf32v5.json
tail -3 gpurun_out/gpu_tests_$tag.log; tail -2 gpurun_out/smoke_$tag.log; cut -c1-400 gpurun_out/bench_$tag.json
