set -x
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/gpu_tests_v18.log
python bench.py > gpurun_out/bench_v18.json 2> gpurun_out/bench_v18.err
cp gpurun_out/op_breakdown.json gpurun_out/op_breakdown_v18.json
tail -8 gpurun_out/gpu_tests_v18.log
cat gpurun_out/bench_v18.json | cut -c1-200
