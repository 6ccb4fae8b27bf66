#!/usr/bin/env bash
set -u
T=${1:-r2i}
mkdir -p gpurun_out
step() { echo "== $1" >&2; }
step "decode probe"; python tools/probe_decode_tail.py 2>/dev/null | grep "^{" | tee gpurun_out/${T}_decode_probe.jsonl
step "engine + bench-shape tests"; timeout 900 python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_engine.py tests/test_gpu_fullsize.py -m gpu -x -q -s > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log; grep "bench-shapes" gpurun_out/${T}_tests.log
step "bench N=1"; timeout 700 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench.err; echo "rc=$?"; python - <<P
import json
d = json.load(open("gpurun_out/${T}_bench_1gpu.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["decode_step"]["us"], d["decode_step"]["in_bench_us"], d["output_check"], d["encoder"]["ms"])
P
