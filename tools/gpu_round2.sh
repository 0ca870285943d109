#!/bin/bash
# Round-2 evidence run on the GPU box:  gpurun --timeout 2400 -- 'bash tools/gpu_round2.sh r2_vNN [quick|full]'
# Everything is reduced to text ON THE BOX: gpurun_out/ comes back only if it stays below 64 MiB (two --set full reports
# of 60 + 40 launches were 250 MB and the whole run came back empty).
set -x
tag=${1:-r2_vXX}
mode=${2:-full}
O=gpurun_out
rm -f $O/parity_$tag.txt
SVX_PARITY_LOG=$O/parity_$tag.txt timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $O/gpu_tests_$tag.log
timeout 300 python __graft_entry__.py smoke > $O/smoke_$tag.log 2>&1
timeout 600 python bench.py > $O/bench_$tag.json 2> $O/bench_$tag.err
cp $O/op_breakdown.json $O/op_breakdown_$tag.json
timeout 600 python bench.py --dtype bf16 --views 5 --no-eager --cpu-seconds 3 > $O/bench_${tag}_bf16v5.json 2> $O/bench_${tag}_bf16v5.err
cp $O/op_breakdown.json $O/op_breakdown_${tag}_bf16v5.json
timeout 600 python bench.py --views 5 --no-eager --cpu-seconds 3 > $O/bench_${tag}_tf32v5.json 2> $O/bench_${tag}_tf32v5.err
cp $O/op_breakdown.json $O/op_breakdown_${tag}_tf32v5.json
if [ "$mode" = "full" ]; then
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${tag}_reference.json 2> $O/bench_${tag}_reference.err
  # the stress run of the attention kernel: graph replays with two other streams loading the copy engines and the SMs
  for cfg in "17 1 tf32" "33 1 tf32" "64 3 tf32" "17 1 bf16" "64 5 bf16"; do
    timeout 300 python tools/stress_winattn.py 300 $cfg 2>&1 | grep -E "^stress|Error|FAILED|stall" >> $O/stress_$tag.txt
  done
  M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
  timeout 900 ncu --metrics $M --clock-control none -c 4400 --csv --log-file $O/launches_$tag.csv \
      python bench.py --steps 2 --warmup 3 --no-eager --cpu-seconds 0 > $O/ncu_launch_$tag.log 2>&1
  python tools/launch_summary.py $O/launches_$tag.csv "python bench.py --steps 2 --warmup 3 --no-eager" > $O/launches_${tag}_summary.txt
  timeout 900 ncu --metrics $M --clock-control none -c 4400 --csv --log-file $O/launches_${tag}_bf16v5.csv \
      python bench.py --dtype bf16 --views 5 --steps 2 --warmup 3 --no-eager --cpu-seconds 0 > $O/ncu_launch_${tag}_bf16v5.log 2>&1
  python tools/launch_summary.py $O/launches_${tag}_bf16v5.csv "python bench.py --dtype bf16 --views 5 --steps 2 --warmup 3 --no-eager" > $O/launches_${tag}_bf16v5_summary.txt
  gzip -f $O/launches_$tag.csv $O/launches_${tag}_bf16v5.csv
  # --set full captures, digested here (tools/ncu_summary.py, tools/ncu_lines.py), reports deleted
  full() {  # name, kernel regex, launch-skip, count, bench args...
    local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
    timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip $skip -c $cnt \
        -o $O/ncu_$name python bench.py "$@" --steps 1 --warmup 3 --no-eager --cpu-seconds 0 > $O/ncu_$name.log 2>&1
    python tools/ncu_summary.py $O/ncu_$name.ncu-rep > $O/ncu_${name}_summary.txt 2>&1
    python tools/ncu_lines.py $O/ncu_$name.ncu-rep 0 40 > $O/ncu_${name}_lines0.txt 2>&1
    python tools/ncu_lines.py $O/ncu_$name.ncu-rep 1 40 > $O/ncu_${name}_lines1.txt 2>&1
    rm -f $O/ncu_$name.ncu-rep
  }
  full ${tag}_winattn "winattn_umma_kernel" 36 4
  full ${tag}_gemm "gemm_tf32_kernel" 450 8
  full ${tag}_mlp_slab "mlp_fused_kernel|conv3_slab_kernel" 30 6
  full ${tag}_bf16v5_winattn "winattn_umma_kernel" 36 2 --dtype bf16 --views 5
  full ${tag}_bf16v5_gemm "gemm_bf16_kernel" 450 6 --dtype bf16 --views 5
fi
rm -f $O/*.ncu-rep
du -sm $O
cat $O/stress_$tag.txt
cut -c1-300 $O/bench_${tag}_bf16v5.json; tail -3 $O/bench_${tag}_bf16v5.err; cut -c1-300 $O/bench_${tag}_tf32v5.json
tail -3 $O/gpu_tests_$tag.log; tail -2 $O/smoke_$tag.log; cut -c1-400 $O/bench_$tag.json
