set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/gpu_tests_v16.log
python bench.py > gpurun_out/bench_v16.json 2> gpurun_out/bench_v16.err
cp gpurun_out/op_breakdown.json gpurun_out/op_breakdown_v16.json
SVX_NO_IM2COL=1 python bench.py > gpurun_out/bench_v16_noim2col.json 2> gpurun_out/bench_v16_noim2col.err
tail -3 gpurun_out/gpu_tests_v16.log
cat gpurun_out/bench_v16.json | cut -c1-300
cat gpurun_out/bench_v16_noim2col.json | cut -c1-300
