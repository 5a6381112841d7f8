set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_iou3d_gpu.py tests/test_iou3d_cpu.py -q -x > gpurun_out/s3p_iou.log 2>&1; tail -25 gpurun_out/s3p_iou.log
