set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_box_masks_gpu.py tests/test_reference_on_gpu.py tests/test_backbone_gpu.py tests/test_parity_gpu.py -m gpu -q -x > gpurun_out/c4_tests.log 2>&1; tail -12 gpurun_out/c4_tests.log
for flag in "" "--no-tile-meta"; do
timeout 300 python tools/step_breakdown.py --reps 10 $flag > gpurun_out/c4_bd$flag.json 2> gpurun_out/c4_bd$flag.err
python - <<PY
import json
d = json.loads(open("gpurun_out/c4_bd$flag.json").read().strip().splitlines()[-1])
print("tile_meta", d["tile_meta"], "conv chain", d["conv_chain_us"], "index chain", d["index_chain_us"], "graph", d["graph_us"])
print([ (s["step"][:22], s["us"]) for s in d["steps"]])
PY
done
timeout 300 python tools/tc_timeline.py > gpurun_out/c4_tc_timeline.json 2> gpurun_out/c4_tc_timeline.err; tail -3 gpurun_out/c4_tc_timeline.err
