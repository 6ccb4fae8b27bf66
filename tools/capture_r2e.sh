#!/usr/bin/env bash
set -u
T=${1:-r2e}
mkdir -p gpurun_out
step() { echo "== $1" >&2; }
step "ops + engine tests"; timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
step "attention variants"
for v in default split reg112 reg112poly4 reg104 merged_poly4; do
  if [ "$v" = default ]; then unset TWB200_LIB; else export TWB200_LIB=$PWD/turbo-whisper-workspace_b200/variants/libtwb200_$v.so; fi
  timeout 120 python tools/bench_kernels.py 24 attention 2>/dev/null | grep '^{' | sed "s/^{/{\"variant\": \"$v\", /" >> gpurun_out/${T}_attn_variants.jsonl
done
unset TWB200_LIB
cut -c1-200 gpurun_out/${T}_attn_variants.jsonl
step "logmel sweep"; timeout 100 python - <<'P' 2>/dev/null | tee gpurun_out/${T}_logmel.jsonl
import sys, os, json
sys.path.insert(0, os.getcwd())
import torch
from turbo_whisper_workspace_b200 import ops
dev = torch.device("cuda:0")
for B in (1, 8, 24, 64, 256):
    pcm = torch.randn(B, 480000, device=dev) * 0.1
    lm = ops.LogMel(dev, B)
    ot = torch.zeros(B, 3002, 128, dtype=torch.bfloat16, device=dev)
    for _ in range(3): lm(pcm, None, out_t=ot, out_t_row_off=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20; e0.record()
    for _ in range(n): lm(pcm, None, out_t=ot, out_t_row_off=1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps({"kernel": "logmel (one launch, bf16 time-major out)", "B": B, "ms": round(ms, 4), "GBps": round(B * 2.688e6 / ms / 1e6, 1)}))
P
step "ncu logmel"; timeout 200 ncu --set full --clock-control none -k regex:logmel -c 1 -o /tmp/${T}_logmel -f python tools/ncu_target.py 24 0 0 > gpurun_out/${T}_ncu_logmel.log 2>&1
ncu -i /tmp/${T}_logmel.ncu-rep --page details --csv 2>/dev/null | grep -E "Duration|Throughput|Bank|Active Warps|Issue|Registers|Shared Memory|Executed Ipc|L1/TEX Hit|Warp Cycles Per Issued" | cut -c1-220 > gpurun_out/${T}_ncu_logmel_details.txt; head -40 gpurun_out/${T}_ncu_logmel_details.txt
