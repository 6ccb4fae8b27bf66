#!/usr/bin/env bash
set -u
T=${1:-r2u}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cross_attn.py -x -q 2>&1 | tail -5 | tee gpurun_out/${T}_cross_test.log
rm -f gpurun_out/${T}_cross_probe.jsonl
timeout 200 python tools/probe_cross_attn.py 2>&1 | grep '^{' | tee -a gpurun_out/${T}_cross_probe.jsonl
for v in xa_k64n6 xa_k128n6c1 xa_k64n12c1 xa_k128n2c3; do
  L=$PWD/turbo-whisper-workspace_b200/variants/libtwb200_$v.so
  TWB200_LIB=$L timeout 300 python -m pytest tests/test_gpu_cross_attn.py -x -q 2>&1 | tail -1
  TWB200_LIB=$L timeout 200 python tools/probe_cross_attn.py 2>&1 | grep '^{' | sed "s/\"stream\"/\"$v\"/" | tee -a gpurun_out/${T}_cross_probe.jsonl
done
TWB200_CROSS_ATTN=scalar TWB200_CROSS_SPLITS=4 timeout 200 python tools/probe_cross_attn.py 2>&1 | grep '^{' | tee -a gpurun_out/${T}_cross_probe.jsonl
