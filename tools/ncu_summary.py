#!/usr/bin/env python
"""Turn ncu output brought back from the GPU box into the tracked summaries under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv "title" "command" > profiles/rN_launches_summary.md
      (CSV of `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...`)
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep "title" "command" > profiles/rN_conv_ncu_full.md
      (report of `ncu --set full --clock-control none --import-source on ...`; needs `ncu` on PATH to read it)
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
]


def short(name):
    name = name.replace("void ", "").replace("btc::", "")
    return name.split("(")[0]


def launches(path, title, command):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        k = short(r[ki])
        tot[k] += float(r[vi].replace(",", "")) / 1e3
        cnt[k] += 1
    total = sum(tot.values())
    print("# %s\n" % title)
    print("Command (B200, under gpurun): `%s`\n(cold-cache, serialised launches: compare SHARES, not absolute times).  Raw list: `%s`.\n"
          % (command, path.split("/")[-1]))
    print("| launches | total us | share | kernel |\n|---:|---:|---:|---|")
    for k, v in tot.most_common():
        print("| %d | %.1f | %.1f%% | `%s` |" % (cnt[k], v, 100 * v / total, k))
    conv = sum(v for k, v in tot.items() if k.startswith("conv_fwd"))
    print("\nTotal profiled: %.1f us over %d launches.  Gather-GEMM conv kernels: %.1f%% of device time."
          % (total, sum(cnt.values()), 100 * conv / total))


def full(path, title, command):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("# %s\n" % title)
    print("Command: `%s`\n" % command)
    ki = hdr.index("Kernel Name")
    print("Launches: " + "; ".join("%d = `%s`" % (i, short(d[ki])) for i, d in enumerate(data)) + "\n")
    print("| metric | " + " | ".join("launch %d" % i for i in range(len(data))) + " | unit |")
    print("|---|" + "---|" * (len(data) + 1))
    for m in FULL_METRICS:
        if m in hdr:
            i = hdr.index(m)
            print("| %s | %s | %s |" % (m, " | ".join(d[i] for d in data), units[i]))


if __name__ == "__main__":
    kind, path, title, command = sys.argv[1:5]
    (launches if kind == "launches" else full)(path, title, command)
