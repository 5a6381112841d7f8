#!/bin/bash
# Round-2 ncu captures (run under gpurun on one B200).  Only CSV exports and the gather-GEMM report travel back
# (gpurun_out/ is capped at 64 MiB); tools/ncu_extract.py / ncu_summary.py turn them into the summaries under profiles/.
set -u
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 1 --no-graph --no-extra --no-cpu-baseline"
IDX='regex:mark|scan_rank|scan_|tables2|subm_table|fill_table|tile_meta|hash_insert|vox_|index_clear'
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size,launch__registers_per_thread
# 1 every launch of the (eager) bench command with its device time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv $BENCH > gpurun_out/r2_launches.log 2>&1
# 2 the twelve gather-GEMM launches of the last timed step (batch 1 = the batch the bench line's roofline is quoted on)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_tc -s 36 -c 12 -f -o gpurun_out/r2_conv $BENCH > gpurun_out/r2_conv.log 2>&1
ncu -i gpurun_out/r2_conv.ncu-rep --page raw --csv > gpurun_out/r2_conv_raw.csv 2>/dev/null
# 3 the indexing kernels of one bench step (batch 16 x 20k points): memory metrics only (a handful of passes)
timeout 600 ncu --metrics $MET --clock-control none -k "$IDX" -s 60 -c 60 --csv --log-file gpurun_out/r2_index_bench.csv $BENCH > gpurun_out/r2_index_bench.log 2>&1
# 4 the indexing kernels on BASELINE config 5 (4 x 100k-point clouds, voxel 0.025 m)
timeout 600 ncu --metrics $MET --clock-control none -k "$IDX" -s 0 -c 90 --csv --log-file gpurun_out/r2_index_stress.csv python tools/stress_bench.py --batch 4 --reps 1 > gpurun_out/r2_index_stress.log 2>&1
ls -la gpurun_out/ | head -30
du -sh gpurun_out
