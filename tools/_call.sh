set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s3n_tests.log 2>&1; tail -3 gpurun_out/s3n_tests.log
timeout 400 python tools/step_breakdown.py --reps 10 > gpurun_out/s3n_bd.json 2> gpurun_out/s3n_bd.err; tail -2 gpurun_out/s3n_bd.err
timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/s3n_bench.json 2> gpurun_out/s3n_bench.err; tail -3 gpurun_out/s3n_bench.err
