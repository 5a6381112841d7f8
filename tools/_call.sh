set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_backbone_gpu.py -q -x > gpurun_out/s4b_tests.log 2>&1; tail -3 gpurun_out/s4b_tests.log
timeout 200 python tools/step_breakdown.py --reps 10 > gpurun_out/s4b_bd.json 2> gpurun_out/s4b_bd.err; tail -2 gpurun_out/s4b_bd.err
