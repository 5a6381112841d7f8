set -u
mkdir -p gpurun_out
timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/s3z_bench.json 2> gpurun_out/s3z_bench.err; tail -3 gpurun_out/s3z_bench.err
timeout 600 python -m pytest tests/test_backbone_gpu.py tests/test_static_mode_gpu.py -q -x > gpurun_out/s3z_tests.log 2>&1; tail -2 gpurun_out/s3z_tests.log
