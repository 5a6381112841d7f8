set -u
mkdir -p gpurun_out
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 120 python -m pytest tests -m gpu -q -x > gpurun_out/tests.log 2>&1; tail -3 gpurun_out/tests.log
