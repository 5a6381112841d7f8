#!/usr/bin/env python
"""Read an `ncu --set full` report brought back from the GPU box and write a compact JSON of the metrics the roofline
discussion uses (needs `ncu` on PATH; runs on the CPU container).

    python tools/ncu_extract.py gpurun_out/r2_conv.ncu-rep "command that produced it" > profiles/r2_conv_tc_ncu.json
"""
import csv
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "lts__t_sectors.sum": "l2_sectors",
    "lts__t_bytes.sum": "l2_bytes",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct_of_peak",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_read_bytes",
    "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1tex_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_hmma_pct_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_active",
    "sm__inst_executed_pipe_tensor.sum": "tensor_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed.sum": "instructions",
    "sm__cycles_elapsed.avg": "sm_cycles_elapsed",
    "sm__cycles_active.avg": "sm_cycles_active",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct": "stall_long_scoreboard_pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct": "stall_barrier_pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct": "stall_mio_throttle_pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct": "stall_lg_throttle_pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct": "stall_short_scoreboard_pct",
}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return x


def main():
    path, command = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    cols = {m: hdr.index(m) for m in WANT if m in hdr}
    res = {"command": command, "units": {WANT[m]: units[i] for m, i in cols.items()}, "launches": []}
    for d in data:
        name = d[ki].replace("void ", "").replace("btc::", "").split("(")[0]
        res["launches"].append(dict({"kernel": name}, **{WANT[m]: num(d[i]) for m, i in cols.items()}))
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
