#!/usr/bin/env bash
set -u
T=${1:-r2s}
mkdir -p gpurun_out
run() {
  label=$1; shift
  timeout 600 python bench.py --warmup 5 --no-extras --no-cpu-baseline "$@" 2>gpurun_out/${T}_err_$label.txt | python -c "
import json,sys
l=sys.stdin.readline()
if not l.strip(): print(json.dumps({'cfg':'$label','error':'no line'})); sys.exit(0)
d=json.loads(l)
print(json.dumps({'cfg':'$label','value':round(d['value'],1),'e2e':round(d['e2e']['value'],1),'ms_per_step':round(d['ms_per_step'],1),'steps':d['steps'],'ok':[d['output_check']['e2e_rows_equal_single_context'],d['output_check']['resident_rows_equal_single_context']]}))" | tee -a gpurun_out/${T}_bench_ab.jsonl
}
for K in 20 8 3; do
run mb96_c3_k$K --steps $K --max-batch 96 --contexts 3
run mb96_c4_k$K --steps $K --max-batch 96 --contexts 4
run mb96_c5_k$K --steps $K --max-batch 96 --contexts 5
run mb72_c4_k$K --steps $K --max-batch 72 --contexts 4
run mb24_c4_k$K --steps $K --max-batch 24 --contexts 4
done
