#!/usr/bin/env bash
# end of round: full GPU suite on the final tree, ncu of both cross-attention kernels at 96 rows, one more context-count point
set -u
T=${1:-r2end}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${T}_gpu_suite_tail.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
PROBE_ROWS=96 TWB200_CROSS_ATTN=stream timeout 300 ncu --set full --clock-control none -k regex:"cross_attn_stream" -s 4 -c 2 -o /tmp/${T}_xs -f python tools/probe_cross_attn.py > gpurun_out/${T}_ncu_stream.log 2>&1
PROBE_ROWS=96 TWB200_CROSS_SPLITS=4 timeout 300 ncu --set full --clock-control none -k regex:"decode_attn" -s 4 -c 2 -o /tmp/${T}_xd -f python tools/probe_cross_attn.py > gpurun_out/${T}_ncu_scalar.log 2>&1
for r in xs xd; do ncu -i /tmp/${T}_$r.ncu-rep --page raw --csv --metrics $M > gpurun_out/${T}_ncu_${r}_raw.csv 2>/dev/null; done
timeout 400 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --contexts 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(json.dumps({'cfg':'mb96_c6_k20','value':round(d['value'],1),'e2e':round(d['e2e']['value'],1),'ok':d['output_check']['resident_rows_equal_single_context']}))" | tee gpurun_out/${T}_c6.jsonl
