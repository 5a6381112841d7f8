set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s3j_tests.log 2>&1; tail -12 gpurun_out/s3j_tests.log
timeout 900 python bench.py > gpurun_out/s3j_bench.json 2> gpurun_out/s3j_bench.err; tail -3 gpurun_out/s3j_bench.err; cat gpurun_out/s3j_bench.json | cut -c1-1500
