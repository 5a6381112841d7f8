set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3x_smoke.log 2>&1; tail -1 gpurun_out/s3x_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s3x_tests.log 2>&1; tail -3 gpurun_out/s3x_tests.log
timeout 900 python bench.py > gpurun_out/s3x_bench.json 2> gpurun_out/s3x_bench.err; tail -3 gpurun_out/s3x_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-extra --no-cpu-baseline > gpurun_out/r2_launches.log 2>&1; tail -1 gpurun_out/r2_launches.log | cut -c1-200
timeout 300 python tools/tc_timeline.py > gpurun_out/s3x_tl.json 2> gpurun_out/s3x_tl.err
timeout 400 python tools/step_breakdown.py --reps 10 --diag > gpurun_out/s3x_bd.json 2> gpurun_out/s3x_bd.err
