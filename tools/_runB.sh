set -x
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/gpu_tests_v19.log
python bench.py > gpurun_out/bench_v19.json 2> gpurun_out/bench_v19.err
cp gpurun_out/op_breakdown.json gpurun_out/op_breakdown_v19.json
tail -5 gpurun_out/gpu_tests_v19.log
tail -3 gpurun_out/bench_v19.err
cat gpurun_out/bench_v19.json | cut -c1-700
