#!/usr/bin/env bash
set -u
T=${1:-r2v}
mkdir -p gpurun_out
run() {
  label=$1; shift
  timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>gpurun_out/${T}_err_$label.txt | python -c "
import json,sys
l=sys.stdin.readline()
if not l.strip(): print(json.dumps({'cfg':'$label','error':'no line'})); sys.exit(0)
d=json.loads(l)
print(json.dumps({'cfg':'$label','value':round(d['value'],1),'e2e':round(d['e2e']['value'],1),'ms_per_step':round(d['ms_per_step'],1),'cross_us':round(d['roofline']['ms_per_launch']*1e3,1),'dec_us':round(d['decode_step']['us'],1),'in_bench_us':round(d['decode_step']['in_bench_us'],1),'ok':[d['output_check']['e2e_rows_equal_single_context'],d['output_check']['resident_rows_equal_single_context']]}))" | tee -a gpurun_out/${T}_bench_ab.jsonl
}
V=$PWD/turbo-whisper-workspace_b200/variants
run stream
TWB200_CROSS_ATTN=scalar TWB200_CROSS_SPLITS=4 run scalar
TWB200_LIB=$V/libtwb200_xa_k128n2c3.so run stream_c3
run stream_again
TWB200_CROSS_ATTN=scalar TWB200_CROSS_SPLITS=4 run scalar_again
