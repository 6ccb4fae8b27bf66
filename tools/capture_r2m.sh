#!/usr/bin/env bash
# r2m: ncu --set full + source-level stalls of the two kernels that fill the GPU in a decode step: LM head and cross-attention
set -u
T=${1:-r2m}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"lmhead|decode_attn" -s 9 -c 9 -o /tmp/${T}_dec -f python tools/ncu_target.py 24 3 0 > gpurun_out/${T}_ncu.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,l1tex__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__occupancy_limit_registers,launch__waves_per_multiprocessor,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__inst_executed.sum,sm__cycles_elapsed.avg.per_second
ncu -i /tmp/${T}_dec.ncu-rep --page raw --csv --metrics $M > gpurun_out/${T}_raw.csv 2>/dev/null
ncu -i /tmp/${T}_dec.ncu-rep --page source --csv > /tmp/${T}_source.csv 2>/dev/null
python - <<P > gpurun_out/${T}_stalls.txt
import csv
text = open("/tmp/${T}_source.csv").read()
# the source page lists kernels one after another; split on header rows
rows = list(csv.reader(text.splitlines()))
blocks, cur = [], None
for r in rows:
    if "Source" in r and "# Samples" in r:
        cur = {"h": r, "body": []}; blocks.append(cur)
    elif cur is not None:
        cur["body"].append(r)
def val(r, i):
    try: return float(r[i])
    except Exception: return 0.0
for bi, b in enumerate(blocks):
    h = b["h"]; si, src = h.index("# Samples"), h.index("Source")
    stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(val(r, si) for r in b["body"])
    print(f"=== kernel block {bi}: {len(b['body'])} lines, {int(tot)} samples")
    agg = {h[i]: sum(val(r, i) for r in b["body"]) for i in stall}
    print("   ", sorted(((k, int(v)) for k, v in agg.items() if v > 0), key=lambda x: -x[1]))
    top = sorted(range(len(b["body"])), key=lambda n: -val(b["body"][n], si))[:28]
    for n in sorted(top):
        r = b["body"][n]
        extra = " ".join(f"{h[i][6:]}={int(val(r, i))}" for i in stall if val(r, i) >= 0.1 * max(1.0, val(r, si)))
        print(f"{n:5d} {r[src][:64]:64s} {int(val(r, si)):6d} {extra}")
P
wc -l gpurun_out/${T}_stalls.txt; du -sh gpurun_out >&2
