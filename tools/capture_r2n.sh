#!/usr/bin/env bash
set -u
T=${1:-r2n}
mkdir -p gpurun_out
run() {  # label, env...
  label=$1; shift
  env "$@" python tools/probe_decode_tail.py 2>/dev/null | grep "^{" | sed "s/^{/{\"cfg\": \"$label\", /" | tee -a gpurun_out/${T}_decode_probe.jsonl
  env "$@" timeout 600 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print(json.dumps({'cfg':'$label','value':round(d['value'],1),'e2e':round(d['e2e']['value'],1),'ms_per_step':round(d['ms_per_step'],1),'decode_alone_us':round(d['decode_step']['us'],1),'in_bench_us':round(d['decode_step']['in_bench_us'],1),'xattn_ms':round(d['roofline']['ms_per_launch'],4),'ok':d['output_check']['e2e_rows_equal_single_context']}))" | tee -a gpurun_out/${T}_bench_ab.jsonl
}
L=$PWD/turbo-whisper-workspace_b200/variants/libtwb200_lmh16.so
run base A=1
run lmh16 TWB200_LIB=$L
run splits7 TWB200_CROSS_SPLITS=7
run splits6 TWB200_CROSS_SPLITS=6
run lmh16_splits7 TWB200_LIB=$L TWB200_CROSS_SPLITS=7
run base2 A=1
