#!/usr/bin/env bash
# r2d: one-launch log-mel on the GPU (tests + timing), whole GPU suite, full stall listing of the attention kernel
set -u
T=${1:-r2d}
mkdir -p gpurun_out
step() { echo "== $1" >&2; }
step "gpu test suite"; timeout 900 python -m pytest tests -m gpu -x -q --durations=8 -s > gpurun_out/${T}_gpu_suite.log 2>&1; tail -4 gpurun_out/${T}_gpu_suite.log
step "kernel micro-benches"; timeout 150 python tools/bench_kernels.py 24 2>/dev/null | grep '^{' > gpurun_out/${T}_kernels.jsonl; grep -E "logmel|layernorm" gpurun_out/${T}_kernels.jsonl
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
python - <<P > gpurun_out/${T}_attn_stalls.txt
import csv
rows = list(csv.reader(open("/tmp/${T}_attn_source.csv")))
hdr = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
h = rows[hdr]; body = rows[hdr + 1:]
si, src = h.index("# Samples"), h.index("Source")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") or c.lower().startswith("stall")]
print("columns:", h)
def val(r, i):
    try: return float(r[i])
    except Exception: return 0.0
tot = sum(val(r, si) for r in body)
print("total samples", tot, "lines", len(body))
cum = 0
for n, r in enumerate(body):
    v = val(r, si)
    if v >= 20:
        extra = " ".join(f"{h[i]}={r[i]}" for i in stall_cols if val(r, i) >= 10)
        print(f"{n:5d} {r[src][:70]:70s} {int(v):6d} {extra}")
P
head -c 3000 gpurun_out/${T}_attn_stalls.txt | head -5
wc -l gpurun_out/${T}_attn_stalls.txt
du -sh gpurun_out >&2
