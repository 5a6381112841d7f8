set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s3f_tests.log 2>&1; tail -3 gpurun_out/s3f_tests.log
timeout 400 python tools/step_breakdown.py --reps 10 > gpurun_out/s3f_bd.json 2> gpurun_out/s3f_bd.err; tail -2 gpurun_out/s3f_bd.err
timeout 300 python tools/tc_timeline.py > gpurun_out/s3f_tl.json 2> gpurun_out/s3f_tl.err; tail -3 gpurun_out/s3f_tl.err
timeout 300 python tools/chain_profile.py > gpurun_out/s3f_chain.json 2> gpurun_out/s3f_chain.err; tail -3 gpurun_out/s3f_chain.err
timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/s3f_bench.json 2> gpurun_out/s3f_bench.err; tail -3 gpurun_out/s3f_bench.err
