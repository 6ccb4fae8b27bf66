#!/usr/bin/env bash
# 8 GPUs: the contract bench under torchrun at N = 8 (headline weak scaling + config3 = ONE 1 h file sharded over 8 ranks)
set -u
T=${1:-r2g8}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/${T}_bench_8gpu.json 2> gpurun_out/${T}_bench_8gpu.err; echo "rc=$?"
python - <<P
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_8gpu.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["value"])
    print("config3:", json.dumps(d.get("config3"))[300:1400])
    print("config4:", d["config4"]["rtfx"], d["config4"]["seconds"])
    print("output_check:", d.get("output_check"))
except Exception as e:
    print("no json", e)
P
tail -5 gpurun_out/${T}_bench_8gpu.err | cut -c1-300
