#!/bin/bash
# One evidence round on the GPU box (run through gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh vNN'
# writes gpurun_out/{gpu_tests,bench,bench_reference,op_breakdown,launches,prof_gemm}_vNN.*; afterwards, in the build container:
#   python tools/launch_summary.py gpurun_out/launches_r1_vNN.csv "<cmd>" > profiles/r1_launches_vNN_summary.txt
#   python tools/ncu_summary.py gpurun_out/prof_gemm_vNN.ncu-rep > profiles/r1_ncu_gemm_vNN.txt
#   python tools/ncu_lines.py gpurun_out/prof_gemm_vNN.ncu-rep 0 25        # stall samples per CUDA source line
#   python tools/op_groups.py gpurun_out/op_breakdown_vNN.json 40          # per-layer times vs the per-op roofline floor
set -x
tag=${1:-vXX}
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/gpu_tests_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
cp gpurun_out/op_breakdown.json gpurun_out/op_breakdown_$tag.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2> gpurun_out/bench_${tag}_reference.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4400 --csv \
    --log-file gpurun_out/launches_r1_$tag.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm_tf32_kernel" --launch-skip 60 -c 6 \
    -o gpurun_out/prof_gemm_$tag python tools/run_module.py encoder 64 3 1 > gpurun_out/ncu_gemm_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"mlp_fused_kernel|conv3_slab_kernel|winattn_kernel" -c 8 \
    -o gpurun_out/prof_other_$tag python tools/run_module.py all 64 3 1 > gpurun_out/ncu_other_$tag.log 2>&1
tail -2 gpurun_out/gpu_tests_$tag.log; cut -c1-300 gpurun_out/bench_$tag.json
