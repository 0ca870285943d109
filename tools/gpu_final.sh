#!/bin/bash
# Final evidence run of a round (≈ 10 min of box time):  gpurun --timeout 780 -- 'bash tools/gpu_final.sh r2_vNN'
# smoke, bench lines (TF32 configs[1], bf16 configs[2], reference arm), GPU tier with the parity log, launch list with DRAM
# bytes, the attention kernel's stress run and its stage-knockout probes.  --set full digests: tools/gpu_round2_lean.sh.
set -x
tag=${1:-r2_vXX}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/box_$tag.txt
timeout 200 python __graft_entry__.py smoke > $O/smoke_$tag.log 2>&1
timeout 300 python bench.py > $O/bench_$tag.json 2> $O/bench_$tag.err
cp $O/op_breakdown.json $O/op_breakdown_$tag.json
rm -f $O/parity_$tag.txt
SVX_PARITY_LOG=$O/parity_$tag.txt timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -25 > $O/gpu_tests_$tag.log
timeout 200 python bench.py --dtype bf16 --views 5 --no-eager --cpu-seconds 3 > $O/bench_${tag}_bf16v5.json 2> $O/bench_${tag}_bf16v5.err
cp $O/op_breakdown.json $O/op_breakdown_${tag}_bf16v5.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none -c 4400 --csv --log-file $O/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-eager --cpu-seconds 0 > $O/ncu_launch_$tag.log 2>&1
python tools/launch_summary.py $O/launches_$tag.csv "python bench.py --steps 2 --warmup 3 --no-eager --cpu-seconds 0" > $O/launches_${tag}_summary.txt
gzip -f $O/launches_$tag.csv
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${tag}_reference.json 2> $O/bench_${tag}_reference.err
{
echo "# tools/stress_winattn.py: encoder forwards as CUDA-graph replays while a second stream floods the copy engines and a"
echo "# third runs unrelated GEMMs; every 25th result compared bit-exact with the first (code $tag)"
for cfg in "64 3 tf32" "64 5 bf16" "17 1 tf32"; do
  timeout 120 python tools/stress_winattn.py 300 $cfg 2>&1 | grep -E "^stress|Error|FAILED|stall"
done
} > $O/winattn_stress_$tag.txt 2>&1
{
echo "# tools/winattn_time.py with the -DSVX_WINATTN_PROBES build: the op alone, 10 back-to-back launches per shape (the H = 7"
echo "# and H = 14 inputs stay L2-resident here); probe bits: 1 no loads, 2 no softmax, 8 no P.V, 16 no stores, 32 no V"
echo "# conversion, 128 strictly ordered MMA issue.  A probed run computes WRONG results by construction (code $tag)"
for pr in 0 1 2 8 16 32 128; do
  SVX_WINATTN_PROBE=$pr SVX_LIB_PATH=swinvox_b200/libswinvox_b200_probes.so timeout 60 python tools/winattn_time.py 192 2>&1 | tail -1
done
SVX_WINATTN_PROBE=0 SVX_LIB_PATH=swinvox_b200/libswinvox_b200_probes.so timeout 60 python tools/winattn_time.py 192 bf16 2>&1 | tail -1
} > $O/winattn_probes_$tag.txt 2>&1
du -sm $O
tail -2 $O/smoke_$tag.log; cut -c1-500 $O/bench_$tag.json; tail -3 $O/bench_$tag.err
cut -c1-300 $O/bench_${tag}_bf16v5.json
head -12 $O/launches_${tag}_summary.txt
tail -4 $O/gpu_tests_$tag.log
cat $O/winattn_stress_$tag.txt $O/winattn_probes_$tag.txt
