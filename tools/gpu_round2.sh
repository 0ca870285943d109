#!/bin/bash
# Round-2 evidence run on the GPU box:  gpurun --timeout 1500 -- 'bash tools/gpu_round2.sh r2_vNN [quick]'
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
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4400 --csv \
      --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-eager > gpurun_out/ncu_launch_$tag.log 2>&1
fi
cut -c1-300 gpurun_out/bench_${tag}_bf16v5.json; tail -3 gpurun_out/bench_${tag}_bf16v5.err; cut -c1-300 gpurun_out/bench_${tag}_tf32v5.json
tail -3 gpurun_out/gpu_tests_$tag.log; tail -2 gpurun_out/smoke_$tag.log; cut -c1-400 gpurun_out/bench_$tag.json
