set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/gpu_tests_v20.log
python bench.py > gpurun_out/bench_v20.json 2> gpurun_out/bench_v20.err
cp gpurun_out/op_breakdown.json gpurun_out/op_breakdown_v20.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_v20_reference.json 2> gpurun_out/bench_v20_reference.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4400 --csv --log-file gpurun_out/launches_r1_v20.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch_v20.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm_tf32_kernel" --launch-skip 60 -c 6 -o gpurun_out/prof_gemm_v20 python tools/run_module.py encoder 64 3 1 > gpurun_out/ncu_gemm_v20.log 2>&1
tail -2 gpurun_out/gpu_tests_v20.log; cut -c1-300 gpurun_out/bench_v20.json; cut -c1-300 gpurun_out/bench_v20_reference.json
