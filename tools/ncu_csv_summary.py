#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics ... --csv --log-file` list (long format: one row per launch and metric):
launches, mean duration, DRAM bytes per launch, achieved DRAM GB/s (= dram bytes / duration) and its fraction of the
measured HBM copy peak, L2 throughput and SM utilisation.

    python tools/ncu_csv_summary.py gpurun_out/r2_index_stress.csv "command" [peak_gbs] > profiles/r2_index_stress_ncu.json
"""
import collections
import csv
import json
import sys


def main():
    path, command = sys.argv[1], sys.argv[2]
    peak = float(sys.argv[3]) if len(sys.argv) > 3 else 6552.0
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 6]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ii, ki, mi, ui, vi = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    launches = collections.OrderedDict()
    for r in rows:
        if r is hdr or not r[ii].isdigit():
            continue
        d = launches.setdefault(r[ii], {"kernel": r[ki].replace("void ", "").replace("btc::", "").split("(")[0]})
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        name = r[mi]
        if name.startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        if name == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)      # -> us
        d[name] = v
    per = collections.OrderedDict()
    for d in launches.values():
        per.setdefault(d["kernel"], []).append(d)
    out = []
    for k, ls in per.items():
        t = sum(x.get("gpu__time_duration.sum", 0.0) for x in ls)
        b = sum(x.get("dram__bytes_read.sum", 0.0) + x.get("dram__bytes_write.sum", 0.0) for x in ls)
        best = max(ls, key=lambda x: (x.get("dram__bytes_read.sum", 0.0) + x.get("dram__bytes_write.sum", 0.0)) /
                   max(x.get("gpu__time_duration.sum", 1e-9), 1e-9))
        bb = best.get("dram__bytes_read.sum", 0.0) + best.get("dram__bytes_write.sum", 0.0)
        bt = best.get("gpu__time_duration.sum", 0.0)
        avg = lambda m: round(sum(x.get(m, 0.0) for x in ls) / len(ls), 2)   # noqa: E731
        out.append({"kernel": k, "launches": len(ls), "mean_us": round(t / len(ls), 2), "dram_MB_per_launch": round(b / len(ls) / 1e6, 3),
                    "dram_GBps": round(b / max(t, 1e-9) / 1e3, 1), "frac_of_hbm_peak": round(b / max(t, 1e-9) / 1e3 / peak, 4),
                    "best_launch": {"us": round(bt, 2), "dram_MB": round(bb / 1e6, 3), "dram_GBps": round(bb / max(bt, 1e-9) / 1e3, 1),
                                    "frac_of_hbm_peak": round(bb / max(bt, 1e-9) / 1e3 / peak, 4)},
                    "dram_pct_of_peak_ncu": avg("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                    "l2_pct_of_peak_ncu": avg("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                    "sm_pct_of_peak_ncu": avg("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                    "grid": int(ls[0].get("launch__grid_size", 0)), "block": int(ls[0].get("launch__block_size", 0)),
                    "registers": int(ls[0].get("launch__registers_per_thread", 0))})
    out.sort(key=lambda r: -r["mean_us"] * r["launches"])
    print(json.dumps({"command": command, "hbm_peak_gbs": peak, "note": "launches are serialised and cold-cache under ncu; dram bytes are "
                      "the kernel's own DRAM traffic (reads + writes) per launch", "kernels": out}, indent=1))


if __name__ == "__main__":
    main()
