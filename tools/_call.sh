set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_backbone_gpu.py -q -x -k pipelined > gpurun_out/s4a_tests.log 2>&1; tail -3 gpurun_out/s4a_tests.log
