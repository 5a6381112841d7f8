#!/usr/bin/env bash
# First GPU call of the next round: the experiments DESIGN.md §8.1(0) asks for, cheapest first (~2 min on one B200).
#   gpurun --timeout 600 -- 'bash tools/round2_battery.sh'
# Everything lands in gpurun_out/battery_*.  Nothing here changes defaults; read the numbers, then decide.
set -u
mkdir -p gpurun_out
# 1. cycles per MMA against MMAs per tcgen05.commit, in isolation (modes 6 / 7 at the end of the output)
[ -x tools/mma_rate ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_rate tools/mma_rate.cu
timeout 60 ./tools/mma_rate > gpurun_out/battery_mma_rate.txt 2>&1 || echo "mma_rate failed" >> gpurun_out/battery_mma_rate.txt
tail -22 gpurun_out/battery_mma_rate.txt
# 2. commit groups: bit-identity with the per-stage-commit tile (a hang is cut by pytest's timeout)
BTC_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_parity_gpu.py -k commit_groups -x -q --timeout 120 \
    > gpurun_out/battery_cg_tests.log 2>&1
tail -3 gpurun_out/battery_cg_tests.log
# 3. step anatomy + timing diagnostics for commit groups 1 / 2 / 3 (only if 2. passed)
if grep -q " passed" gpurun_out/battery_cg_tests.log && ! grep -q "failed\|error" gpurun_out/battery_cg_tests.log; then
    for cg in 1 2 3; do
        timeout 120 python tools/step_breakdown.py --commit-group $cg --diag --reps 10 > gpurun_out/battery_breakdown_cg$cg.json \
            2> gpurun_out/battery_breakdown_cg$cg.err
        python - <<PY
import json
d = json.loads(open("gpurun_out/battery_breakdown_cg$cg.json").read().strip().splitlines()[-1])
print("commit group $cg: conv chain %.0f us, graph %.0f us" % (d["conv_chain_us"], d["graph_us"]))
for r in d["diag"]:
    print("   mask %2d %-40s conv chain %7.1f  32->32 %6.1f  64->64 %6.1f" % (r["mask"], r["what"], r["conv_chain_us"], r["conv32_us"], r["conv64_us"]))
PY
    done
fi
# 4. step level: rulebook chain of batch i+1 overlapped with the conv chain of batch i (engine.OverlappedBackbone)
BTC_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_backbone_gpu.py -k overlapped -x -q --timeout 150 \
    > gpurun_out/battery_overlap_test.log 2>&1
tail -3 gpurun_out/battery_overlap_test.log
if grep -q " passed" gpurun_out/battery_overlap_test.log; then
    timeout 200 python bench.py --overlap --no-cpu-baseline > gpurun_out/battery_bench_overlap.log 2> gpurun_out/battery_bench_overlap.err
    python - <<PY
import json
b = json.loads(open("gpurun_out/battery_bench_overlap.log").read().strip().splitlines()[-1])
print("value", b["value"], "e2e", b["e2e"]["value"], "e2e_overlapped", b.get("e2e_overlapped"))
PY
fi
# 5. 8 vs 16 producer warps at the current state of the tile
timeout 150 python tools/step_breakdown.py --variants "16,0,1;8,0,1" --reps 10 > gpurun_out/battery_npw.json 2> gpurun_out/battery_npw.err
python - <<PY
import json
for line in open("gpurun_out/battery_npw.json"):
    d = json.loads(line)
    print(d["variant"], "conv chain", d["conv_chain_us"], "graph", d["graph_us"])
PY
# 6. training step (config 4 shape, eager shim: fwd + bwd + Adam) and the mask-path timings, for profiles/
timeout 200 python tools/train_step.py --steps 10 > gpurun_out/battery_train_step.json 2> gpurun_out/battery_train_step.err
tail -1 gpurun_out/battery_train_step.json
timeout 120 python -m pytest tests/test_mask_timings_gpu.py tests/test_stress_properties_gpu.py -m gpu -x -q > gpurun_out/battery_masks_stress.log 2>&1
tail -3 gpurun_out/battery_masks_stress.log
# 7. programmatic dependent launch of the conv tile (prologue of layer l+1 under the tail of layer l)
BTC_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_parity_gpu.py -k pdl_chain -x -q --timeout 120 > gpurun_out/battery_pdl_test.log 2>&1
tail -3 gpurun_out/battery_pdl_test.log
if grep -q " passed" gpurun_out/battery_pdl_test.log; then
    timeout 120 python tools/step_breakdown.py --pdl --reps 10 > gpurun_out/battery_breakdown_pdl.json 2> gpurun_out/battery_breakdown_pdl.err
    python - <<PY
import json
d = json.loads(open("gpurun_out/battery_breakdown_pdl.json").read().strip().splitlines()[-1])
print("PDL: conv chain %.0f us, graph %.0f us" % (d["conv_chain_us"], d["graph_us"]))
PY
fi
