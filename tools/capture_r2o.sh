#!/usr/bin/env bash
set -u
T=${1:-r2o}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?"
python - <<P
import json
d = json.load(open("gpurun_out/${T}_bench.json"))
print(d["value"], d["config2_beams5"], d["config4"]["rtfx"], d["config3"]["30_5"]["rtfx"])
P
tail -3 gpurun_out/${T}_bench.err | cut -c1-300
