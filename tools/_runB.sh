set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/gpu_tests_v17.log
python bench.py > gpurun_out/bench_v17.json 2> gpurun_out/bench_v17.err
cp gpurun_out/op_breakdown.json gpurun_out/op_breakdown_v17.json
tail -3 gpurun_out/gpu_tests_v17.log
cat gpurun_out/bench_v17.json | cut -c1-200
# captures: swin stage-0 block (norm1, qkv, attn, proj, norm2, fc1, fc2) = the first lnrows launch after patch-embed onwards
ncu --set full --clock-control none --import-source on -k regex:"gemm_tf32_kernel|winattn|lnrows" --launch-skip 45 -c 12 -o gpurun_out/prof_swin0_v17 python tools/run_module.py encoder 64 3 1 > gpurun_out/ncu_swin0.log 2>&1
SVX_ISOLATE=1 ncu --set full --clock-control none --import-source on -k regex:"conv3_slab" -c 7 -o gpurun_out/prof_merger_v17 python tools/run_module.py merger 64 3 1 > gpurun_out/ncu_merger.log 2>&1
