#!/usr/bin/env bash
# last checks of the round: reference arm incl. the literal num_beams = 5 call, and the own arm's line after the traffic-key change
set -u
T=${1:-r2zz}
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo rc=$?
cut -c1-200 gpurun_out/${T}_bench_reference.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/${T}_bench_1gpu_noextras.json 2> gpurun_out/${T}_bench.err; echo rc=$?
python - <<P
import json
d=json.loads(open("gpurun_out/${T}_bench_reference.json").readline()); print(d["value"], d["ms_per_step"], d["literal_num_beams_5"])
d=json.loads(open("gpurun_out/${T}_bench_1gpu_noextras.json").readline()); r=d["roofline"]
print(d["value"], d["e2e"]["value"], {k:r[k] for k in ("achieved","frac","bytes_per_launch","traffic","rows_per_launch")}, r["at_24_rows"])
P
