set -u
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_roi_pool_gpu.py -q --tb=short > gpurun_out/n1_tests.log 2>&1; tail -40 gpurun_out/n1_tests.log
timeout 90 python tools/roi_pool_report.py gpurun_out/roi_pool_report.json > gpurun_out/n1_report.log 2>&1; tail -c 1500 gpurun_out/n1_report.log
