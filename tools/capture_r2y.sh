#!/usr/bin/env bash
# ncu of one decode step at 96 rows: where do the skinny GEMMs spend their time (L2 -> SM activations traffic?)
set -u
T=${1:-r2y}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:"skinny|lmhead|decode_attn|finalize|embed" -s 37 -c 37 -o /tmp/${T}_dec96 -f python tools/ncu_target.py 96 3 0 > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_bytes.sum,l1tex__t_sector_hit_rate.pct,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__waves_per_multiprocessor,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
ncu -i /tmp/${T}_dec96.ncu-rep --page raw --csv --metrics $M > gpurun_out/${T}_ncu_dec96_raw.csv 2>/dev/null
python - <<P
import csv
rows=list(csv.reader(open("gpurun_out/${T}_ncu_dec96_raw.csv")))
h=rows[0]; ki=h.index("Kernel Name")
def col(n): return h.index(n)
for r in rows[2:]:
    print(r[ki][:46].ljust(46), "us", r[col("gpu__time_duration.sum")], "grid", r[col("launch__grid_size")], "dramR", r[col("dram__bytes_read.sum")], "lts", r[col("lts__t_bytes.sum")], "lts%", r[col("lts__throughput.avg.pct_of_peak_sustained_elapsed")][:5], "l1B", r[col("l1tex__t_bytes.sum")], "l1hit", r[col("l1tex__t_sector_hit_rate.pct")][:5], "issue%", r[col("smsp__issue_active.avg.pct_of_peak_sustained_active")][:5])
P
