#!/usr/bin/env bash
# N-GPU validation of the bench defaults under torchrun (driver's launch line): bash tools/capture_r2t.sh <tag> [N] [extras]
set -u
T=${1:-r2t}; N=${2:-2}; X=${3:-config3,config2_beams5,config4}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 --extras $X > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err
echo "rc=$?"
python - <<P
import json
d=json.loads(open("gpurun_out/${T}_bench_${N}gpu.json").readline())
print({k:d[k] for k in ("value","n_gpus","ms_per_step","output_check")}, d["e2e"]["value"])
for k in ("config3","config2_beams5","config4"):
    if k in d: print(k, json.dumps(d[k])[:700])
P
tail -3 gpurun_out/${T}_bench_${N}gpu.err | cut -c1-300
