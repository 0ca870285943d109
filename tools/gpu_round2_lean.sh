#!/bin/bash
# Lean round-2 evidence run (≈ 22 min of box time):  gpurun --timeout 1500 -- 'bash tools/gpu_round2_lean.sh r2_vNN'
# Order = value of the evidence: smoke, bench lines, launch list, --set full digests of the two top kernels, then the
# GPU test tier with the parity log.  Every ncu report is digested ON THE BOX and deleted (gpurun_out/ returns only
# below 64 MiB).  Numbers printed under ncu are never bench values: the bench lines come from the plain runs above them.
set -x
tag=${1:-r2_vXX}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/box_$tag.txt
timeout 300 python __graft_entry__.py smoke > $O/smoke_$tag.log 2>&1
timeout 420 python bench.py > $O/bench_$tag.json 2> $O/bench_$tag.err
cp $O/op_breakdown.json $O/op_breakdown_$tag.json
timeout 300 python bench.py --dtype bf16 --views 5 --no-eager --cpu-seconds 3 > $O/bench_${tag}_bf16v5.json 2> $O/bench_${tag}_bf16v5.err
cp $O/op_breakdown.json $O/op_breakdown_${tag}_bf16v5.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 420 ncu --metrics $M --clock-control none -c 4400 --csv --log-file $O/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-eager --cpu-seconds 0 > $O/ncu_launch_$tag.log 2>&1
python tools/launch_summary.py $O/launches_$tag.csv "python bench.py --steps 2 --warmup 3 --no-eager --cpu-seconds 0" > $O/launches_${tag}_summary.txt
gzip -f $O/launches_$tag.csv
full() {  # name, kernel regex, launch-skip, count, bench args...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip $skip -c $cnt \
      -o $O/ncu_$name python bench.py "$@" --steps 1 --warmup 3 --no-eager --cpu-seconds 0 > $O/ncu_$name.log 2>&1
  python tools/ncu_summary.py $O/ncu_$name.ncu-rep > $O/ncu_${name}_summary.txt 2>&1
  python tools/ncu_lines.py $O/ncu_$name.ncu-rep 0 40 > $O/ncu_${name}_lines0.txt 2>&1
  rm -f $O/ncu_$name.ncu-rep
}
full ${tag}_gemm "gemm_tf32_kernel" 450 6
full ${tag}_winattn "winattn_umma_kernel" 36 3
rm -f $O/parity_$tag.txt
SVX_PARITY_LOG=$O/parity_$tag.txt timeout 840 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -25 > $O/gpu_tests_$tag.log
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${tag}_reference.json 2> $O/bench_${tag}_reference.err
rm -f $O/*.ncu-rep
du -sm $O
tail -2 $O/smoke_$tag.log; cut -c1-600 $O/bench_$tag.json; tail -3 $O/bench_$tag.err
cut -c1-300 $O/bench_${tag}_bf16v5.json; tail -3 $O/bench_${tag}_bf16v5.err
head -12 $O/launches_${tag}_summary.txt
tail -6 $O/gpu_tests_$tag.log
