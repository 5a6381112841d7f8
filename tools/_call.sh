set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c16_tests.log 2>&1; tail -4 gpurun_out/c16_tests.log
timeout 300 python tools/kernel_times.py --no-tile-meta > gpurun_out/c16_kt.json 2> gpurun_out/c16_kt.err; tail -3 gpurun_out/c16_kt.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/c16_kt.json"))
print("sum", d["sum_us"])
for r in d["kernels"][:14]: print("%-72s %4d %9.1f %8.2f" % (r["kernel"], r["launches"], r["total_us"], r["mean_us"]))
PY
for flag in "" "--no-tile-meta"; do
timeout 300 python tools/step_breakdown.py --reps 10 $flag > "gpurun_out/c16_bd$flag.json" 2> "gpurun_out/c16_bd$flag.err"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/c16_bd$flag.json").read().strip().splitlines()[-1])
    print("flags [$flag] conv chain", d["conv_chain_us"], "index chain", d["index_chain_us"], "graph", d["graph_us"])
    print([s["us"] for s in d["steps"] if s["step"].startswith("conv ")])
except Exception as e:
    print("failed [$flag]", e)
PY
done
timeout 300 python tools/tc_timeline.py > gpurun_out/c16_tc_timeline.json 2> gpurun_out/c16_tc_timeline.err
