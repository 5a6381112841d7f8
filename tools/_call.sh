set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_iou3d_gpu.py > gpurun_out/s3q_tests.log 2>&1; tail -3 gpurun_out/s3q_tests.log
timeout 400 python tools/step_breakdown.py --reps 10 > gpurun_out/s3q_bd.json 2> gpurun_out/s3q_bd.err; tail -2 gpurun_out/s3q_bd.err
timeout 300 python tools/tc_timeline.py > gpurun_out/s3q_tl.json 2> gpurun_out/s3q_tl.err
