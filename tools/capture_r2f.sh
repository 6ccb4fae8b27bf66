#!/usr/bin/env bash
# r2f (2 GPUs): the contract bench under torchrun at N = 2 (config3 sharded by the scheduler), the whole GPU suite with the
# polynomial-exp attention build (candidate default), the multi-GPU single-process sharding test
set -u
T=${1:-r2f}
mkdir -p gpurun_out
step() { echo "== $1" >&2; }
step "bench N=2 (torchrun)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err; echo "rc=$?"; cut -c1-300 gpurun_out/${T}_bench_2gpu.json; tail -5 gpurun_out/${T}_bench_2gpu.err
python - <<P
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_2gpu.json"))
    print("config3:", json.dumps(d.get("config3"))[:900])
    print("config4:", json.dumps(d.get("config4"))[:400])
    print("output_check:", d.get("output_check"))
except Exception as e:
    print("no json", e)
P
step "gpu suite with poly4 attention"
TWB200_LIB=$PWD/turbo-whisper-workspace_b200/variants/libtwb200_poly4.so timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_suite_poly4.log 2>&1; tail -4 gpurun_out/${T}_suite_poly4.log
