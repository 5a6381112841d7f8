set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backbone_gpu.py tests/test_parity_gpu.py tests/test_stress_properties_gpu.py tests/test_chain_gpu.py -m gpu -q > gpurun_out/c9_tests.log 2>&1; tail -25 gpurun_out/c9_tests.log
timeout 300 python tools/kernel_times.py > gpurun_out/c9_kt.json 2> gpurun_out/c9_kt.err; tail -3 gpurun_out/c9_kt.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/c9_kt.json"))
print("sum", d["sum_us"])
for r in d["kernels"][:40]: print("%-72s %4d %9.1f %8.2f" % (r["kernel"], r["launches"], r["total_us"], r["mean_us"]))
PY
for flag in "" "--no-tile-meta"; do
timeout 300 python tools/step_breakdown.py --reps 10 $flag > gpurun_out/c9_bd$flag.json 2> gpurun_out/c9_bd$flag.err
python - <<PY
import json
d = json.loads(open("gpurun_out/c9_bd$flag.json").read().strip().splitlines()[-1])
print("tile_meta", d["tile_meta"], "conv chain", d["conv_chain_us"], "index chain", d["index_chain_us"], "graph", d["graph_us"])
PY
done
