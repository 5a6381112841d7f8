set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3r_smoke.log 2>&1; tail -2 gpurun_out/s3r_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s3r_tests.log 2>&1; tail -3 gpurun_out/s3r_tests.log
timeout 900 python bench.py > gpurun_out/s3r_bench.json 2> gpurun_out/s3r_bench.err; tail -3 gpurun_out/s3r_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s3r_ref.json 2> gpurun_out/s3r_ref.err; cut -c1-300 gpurun_out/s3r_ref.json
