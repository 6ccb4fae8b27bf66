#!/usr/bin/env bash
# the contract's no-flag invocation: must default to N = 1 and finish within minutes
set -u
T=${1:-r2noflags}
mkdir -p gpurun_out
SECONDS=0
timeout 400 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$? wall=${SECONDS}s"
python - <<P
import json
d=json.loads(open("gpurun_out/${T}_bench.json").readline())
print({k:d[k] for k in ("value","steps","warmup","ms_per_step","n_gpus","gpu_launches")}, d["e2e"]["value"], d["output_check"], d["cpu_baseline"]["value"], sorted(k for k in d if k.startswith("config")))
P
