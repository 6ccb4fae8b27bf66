#!/usr/bin/env bash
# A/B of a decode.cu change: shape parity tests, then the bench with the current library and with a variant library
set -u
T=${1:-r2x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_engine.py -x -q 2>&1 | tail -4 | tee gpurun_out/${T}_tests.log
run() {
  label=$1; shift
  timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>gpurun_out/${T}_err_$label.txt | python -c "
import json,sys
l=sys.stdin.readline()
if not l.strip(): print(json.dumps({'cfg':'$label','error':'no line'})); sys.exit(0)
d=json.loads(l)
a=d['decode_step']['at_call_rows']
print(json.dumps({'cfg':'$label','value':round(d['value'],1),'e2e':round(d['e2e']['value'],1),'ms_per_step':round(d['ms_per_step'],1),'cross_us':round(d['roofline']['ms_per_launch']*1e3,1),'cross_frac':round(d['roofline']['frac'],3),'dec24_us':round(d['decode_step']['us'],1),'decR_us':round(a['us'],1),'rows':a['rows'],'in_bench_us':round(d['decode_step']['in_bench_us'],1),'ok':[d['output_check']['e2e_rows_equal_single_context'],d['output_check']['resident_rows_equal_single_context']]}))" | tee -a gpurun_out/${T}_bench_ab.jsonl
}
V=$PWD/turbo-whisper-workspace_b200/variants
run new
TWB200_LIB=$V/libtwb200_decode_prev.so run prev
run new_again
TWB200_LIB=$V/libtwb200_decode_prev.so run prev_again
