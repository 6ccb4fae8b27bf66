#!/usr/bin/env bash
set -u
T=${1:-r2w}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cross_attn.py -x -q 2>&1 | tail -4 | tee gpurun_out/${T}_cross_test.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo rc=$?
python - <<P
import json
d=json.loads(open("gpurun_out/${T}_bench.json").readline())
print(d["value"], d["e2e"]["value"], d["output_check"])
print(json.dumps(d["roofline"])[:1500]); print(json.dumps(d["decode_step"])[:1200])
P
