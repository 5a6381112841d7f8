"""bench.py host logic that can be checked without a GPU: the `roofline` object assembly."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _layers():
    rec = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_b16.json")))
    per_layer = rec["roofline"]["per_layer"]
    specs = []
    n_in = rec["config"]["level_sites"][0]
    for pl in per_layer:
        K = 3 if pl["cout"] == 128 else 27
        specs.append({"n_in": n_in, "n_out": pl["n_out"], "pairs": pl["pairs"], "K": K, "cin": pl["cin"], "cout": pl["cout"],
                      "bytes": 4 * n_in * pl["cin"] + 4 * pl["n_out"] * pl["cout"] + 8 * pl["pairs"] + 4 * K * pl["cin"] * pl["cout"],
                      "flops": 2 * pl["pairs"] * pl["cin"] * pl["cout"]})
        n_in = pl["n_out"]
    return rec, per_layer, specs


def test_dominant_roofline_object():
    import bench
    rec, per_layer, specs = _layers()
    tot_ms = sum(pl["us"] for pl in per_layer) * 1e-3
    roof = bench.dominant_roofline(per_layer, specs, tot_ms, rec["config"]["scenes_per_step_per_gpu"])
    json.dumps(roof)                                        # serialisable
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in roof
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s"
    tr = json.load(open(os.path.join(ROOT, "profiles", "r2_conv_tc_traffic.json")))
    n_out, pairs = tr["config"]["n_out"], tr["config"]["pairs"]
    assert "64->64" in roof["kernel"] and str(n_out) in roof["kernel"]      # the level-3 SubM pair dominates
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-4
    # algorithmic bytes per launch of that layer and the matching ncu capture (same batch, same row count)
    assert roof["alg_bytes_per_launch"] == 4 * n_out * 64 * 2 + 8 * pairs + 4 * 27 * 64 * 64
    assert roof["traffic"] == tr["dram_bytes_per_launch"]
    assert roof["tensor_pipe_pct_active_ncu"] == tr["tensor_pipe_pct_active"]
    assert roof["all_conv_layers"]["launches"] == 12
    # a different workload must not inherit the captured traffic
    roof2 = bench.dominant_roofline(per_layer, specs, tot_ms, 8)
    assert roof2["traffic"] is None
