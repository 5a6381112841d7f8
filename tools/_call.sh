set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_iou3d_gpu.py > gpurun_out/s3w_tests.log 2>&1; tail -3 gpurun_out/s3w_tests.log
timeout 400 python tools/step_breakdown.py --reps 10 > gpurun_out/s3w_bd.json 2> gpurun_out/s3w_bd.err; tail -2 gpurun_out/s3w_bd.err
timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/s3w_bench.json 2> gpurun_out/s3w_bench.err; tail -3 gpurun_out/s3w_bench.err
