set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s3l_tests.log 2>&1; tail -3 gpurun_out/s3l_tests.log
timeout 400 python tools/step_breakdown.py --reps 10 > gpurun_out/s3l_bd.json 2> gpurun_out/s3l_bd.err; tail -2 gpurun_out/s3l_bd.err
timeout 300 python tools/tc_timeline.py > gpurun_out/s3l_tl.json 2> gpurun_out/s3l_tl.err
bash tools/ncu_r2.sh
