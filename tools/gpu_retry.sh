#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> <logfile>   — runs tools/_call.sh on a GPU box, retrying while the pod is busy
T=${1:-1800}; LOG=${2:-/tmp/gpurun.log}
for i in $(seq 1 40); do
  gpurun --timeout "$T" -- 'bash tools/_call.sh' > "$LOG" 2>&1
  if grep -q "status=transient" "$LOG"; then sleep 90; continue; fi
  break
done
