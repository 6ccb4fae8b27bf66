#!/usr/bin/env bash
set -u
T=${1:-r2r}
mkdir -p gpurun_out
echo "== gpu suite" >&2
timeout 1200 python -m pytest tests -m gpu -x -q -s --durations=8 > gpurun_out/${T}_gpu_suite.log 2>&1; tail -12 gpurun_out/${T}_gpu_suite.log | cut -c1-200; grep "bench-shapes" gpurun_out/${T}_gpu_suite.log
echo "== bench (new defaults, with extras)" >&2
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?"
python - <<P
import json
d = json.load(open("gpurun_out/${T}_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "steps")}, "e2e", d["e2e"]["value"], d["detail"]["contexts_per_gpu"], d["decode_step"], d["output_check"])
print("config3", d["config3"]["30_5"], d["config3"]["60_5_literal"]["rtfx"])
print("beams", d["config2_beams5"]["rtfx"], d["config2_beams5"]["microbatches"], "config4", d["config4"]["rtfx"])
P
tail -3 gpurun_out/${T}_bench.err | cut -c1-300
echo "== bench old config for comparison" >&2
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --max-batch 24 --contexts 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('mb24_c4', round(d['value'],1), round(d['e2e']['value'],1))"
