# Scratch command file for `gpurun -- 'bash tools/_call.sh'` (tools/gpu_retry.sh retries while the pod is busy).
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/tests.log 2>&1; tail -3 gpurun_out/tests.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
