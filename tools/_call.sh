set -u
mkdir -p gpurun_out
timeout 300 python tools/tc_timeline.py > gpurun_out/s3b_tl.json 2> gpurun_out/s3b_tl.err; tail -3 gpurun_out/s3b_tl.err
timeout 400 python tools/step_breakdown.py --reps 10 --diag > gpurun_out/s3b_bd.json 2> gpurun_out/s3b_bd.err; tail -2 gpurun_out/s3b_bd.err
