#!/usr/bin/env bash
set -u
T=${1:-r2x3}
mkdir -p gpurun_out
run() {
  label=$1; shift
  timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>gpurun_out/${T}_err_$label.txt | python -c "
import json,sys
l=sys.stdin.readline()
if not l.strip(): print(json.dumps({'cfg':'$label','error':'no line'})); sys.exit(0)
d=json.loads(l)
a=d['decode_step']['at_call_rows']
print(json.dumps({'cfg':'$label','value':round(d['value'],1),'e2e':round(d['e2e']['value'],1),'ms_per_step':round(d['ms_per_step'],1),'decR_us':round(a['us'],1),'rows':a['rows'],'in_bench_us':round(d['decode_step']['in_bench_us'],1),'ok':[d['output_check']['e2e_rows_equal_single_context'],d['output_check']['resident_rows_equal_single_context']]}))" | tee -a gpurun_out/${T}_bench_ab.jsonl
}
TWB200_SKINNY_CHUNK_GRID=2 run selective
run loop
TWB200_SKINNY_CHUNK_GRID=2 run selective_again
run loop_again
