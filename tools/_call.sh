set -u
mkdir -p gpurun_out
timeout 600 python tools/step_breakdown.py --reps 10 --early 3:116,3:88,3:72,2:116,2:88,4:116,5:116,5:100 > gpurun_out/s3u_early.json 2> gpurun_out/s3u_early.err; tail -2 gpurun_out/s3u_early.err; cat gpurun_out/s3u_early.json
