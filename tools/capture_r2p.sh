#!/usr/bin/env bash
# r2p: decode rows beyond 32 per context (one weight stream for two micro-batches): probe + bench A/B
set -u
T=${1:-r2p}
mkdir -p gpurun_out
PROBE_B48=1 python tools/probe_decode_tail.py 2>&1 | grep -E "^\{|rror" | tee gpurun_out/${T}_decode_probe.jsonl
run() {
  label=$1; shift
  timeout 600 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline "$@" 2>gpurun_out/${T}_err_$label.txt | python -c "
import json,sys
l=sys.stdin.readline()
if not l.strip(): print(json.dumps({'cfg':'$label','error':'no line'})); sys.exit(0)
d=json.loads(l)
print(json.dumps({'cfg':'$label','value':round(d['value'],1),'e2e':round(d['e2e']['value'],1),'ms_per_step':round(d['ms_per_step'],1),'in_bench_us':round(d['decode_step']['in_bench_us'],1),'ok':d['output_check']}))" | tee -a gpurun_out/${T}_bench_ab.jsonl
  tail -2 gpurun_out/${T}_err_$label.txt | cut -c1-300
}
run mb24_c4 --max-batch 24 --contexts 4
run mb48_c2 --max-batch 48 --contexts 2
run mb48_c3 --max-batch 48 --contexts 3
run mb48_c4 --max-batch 48 --contexts 4
run mb24_c4b --max-batch 24 --contexts 4
