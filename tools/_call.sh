set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/s3y_topo.txt 2>&1
python -c "import os; print('affinity', len(os.sched_getaffinity(0)), 'cpu_count', os.cpu_count())" >> gpurun_out/s3y_topo.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/s3y_bench8.json 2> gpurun_out/s3y_bench8.err; tail -3 gpurun_out/s3y_bench8.err | cut -c1-300; cut -c1-300 gpurun_out/s3y_bench8.json
