#!/usr/bin/env bash
set -u
T=${1:-r2h}
mkdir -p gpurun_out
step() { echo "== $1" >&2; }
step "gpu test suite"; timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/${T}_gpu_suite.log 2>&1; tail -4 gpurun_out/${T}_gpu_suite.log; grep "bench-shapes\|\[word\]" gpurun_out/${T}_gpu_suite.log | head
step "decode probe"; python tools/probe_decode_tail.py 2>/dev/null | grep "^{" | tee gpurun_out/${T}_decode_probe.jsonl
step "bench N=1"; timeout 700 python bench.py --steps 8 --warmup 3 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench.err; echo "rc=$?"; python - <<P
import json
d = json.load(open("gpurun_out/${T}_bench_1gpu.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["decode_step"], d["output_check"], d["config4"]["rtfx"], d["config4"]["decode_step"]["us"], d["config3"]["30_5"]["rtfx"], d["encoder"]["ms"])
P
