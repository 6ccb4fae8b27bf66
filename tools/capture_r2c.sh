#!/usr/bin/env bash
# r2c: native beam step on the GPU (tests + throughput) and a source-level stall profile of the attention kernel
set -u
T=${1:-r2c}
mkdir -p gpurun_out
step() { echo "== $1" >&2; }
step "beam tests"; timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_pipeline.py tests/test_gpu_zz_word_timestamps.py -m gpu -x -q -s -k "beam or pipeline" > gpurun_out/${T}_beam_tests.log 2>&1; tail -5 gpurun_out/${T}_beam_tests.log
step "beam throughput"; timeout 300 python tools/bench_beams.py > gpurun_out/${T}_beams.jsonl 2> gpurun_out/${T}_beams.err; cat gpurun_out/${T}_beams.jsonl; tail -3 gpurun_out/${T}_beams.err
step "microbatch probe"; timeout 400 python tools/probe_microbatch.py > gpurun_out/${T}_microbatch.jsonl 2>gpurun_out/${T}_microbatch.err; cat gpurun_out/${T}_microbatch.jsonl
step "ncu attention source"
cat > /tmp/attn_only.py <<'P'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from turbo_whisper_workspace_b200 import ops
B, M = 24, 24 * 1500
qkv = torch.randn(M, 3840, device="cuda").to(torch.bfloat16); qkv[:, :1280] *= 0.35
out = torch.empty(M, 1280, dtype=torch.bfloat16, device="cuda")
for _ in range(3): ops.attention_enc(qkv, B, 1500, 20, out=out)
torch.cuda.synchronize()
P
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_enc -s 2 -c 1 -o /tmp/${T}_attn -f python /tmp/attn_only.py > gpurun_out/${T}_ncu_attn.log 2>&1
ncu -i /tmp/${T}_attn.ncu-rep --page source --csv > /tmp/${T}_attn_source.csv 2>/dev/null
python - <<P
import csv, sys
rows = list(csv.reader(open("/tmp/${T}_attn_source.csv")))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and any("Sampl" in c for c in r):
        hdr = i; break
if hdr is None:
    print("no header", rows[:3]); sys.exit(0)
h = rows[hdr]
si = [i for i, c in enumerate(h) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or "Samples" in c]
print("columns:", h)
body = rows[hdr + 1:]
key = si[0] if si else None
def val(r):
    try: return float(r[key])
    except Exception: return 0.0
tot = sum(val(r) for r in body)
print("total samples", tot)
top = sorted(body, key=val, reverse=True)[:70]
for r in top:
    print(" | ".join(c[:60] for c in r[:min(len(r), 12)]))
P
