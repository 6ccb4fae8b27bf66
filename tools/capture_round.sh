#!/usr/bin/env bash
# One-call evidence capture for a round (run under gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/capture_round.sh r2a'
# Writes everything under gpurun_out/<tag>_*; copy what should be judged into profiles/.
# Order: parity first, then the contract bench (its extras carry configs 3 / 4 / 5), then the profiler passes
# (numbers printed by a run under ncu are never bench values).
set -u
T=${1:-rX}
mkdir -p gpurun_out
step() { echo "== $1" >&2; }

step "gpu test suite";        timeout 900 python -m pytest tests -m gpu -x -q --durations=12 -s > gpurun_out/${T}_gpu_suite.log 2>&1; tail -3 gpurun_out/${T}_gpu_suite.log
step "bench N=1";             timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench.err; echo "rc=$?"; cut -c1-400 gpurun_out/${T}_bench_1gpu.json
step "kernel micro-benches";  timeout 150 python tools/bench_kernels.py 24 2>/dev/null | grep '^{' > gpurun_out/${T}_kernels.jsonl; cat gpurun_out/${T}_kernels.jsonl
step "ncu launch list";       timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
[ -f tools/summarize_launches.py ] && python tools/summarize_launches.py gpurun_out/${T}_launches_ncu.csv > gpurun_out/${T}_launches_summary.md 2>/dev/null
step "ncu --set full (decode kernels + one GEMM per shape)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"decode_attn|lmhead|skinny" -c 40 -o gpurun_out/${T}_decode -f python tools/ncu_target.py 24 3 0 > gpurun_out/${T}_ncu_decode.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16|attention_enc|layernorm_kernel|logmel" -c 24 -o gpurun_out/${T}_encoder -f python tools/ncu_target.py 24 0 1 > gpurun_out/${T}_ncu_encoder.log 2>&1
python tools/ncu_traffic.py gpurun_out/${T}_decode.ncu-rep gpurun_out/${T}_encoder.ncu-rep > gpurun_out/${T}_ncu_traffic.json 2>gpurun_out/${T}_ncu_traffic.err
ls -la gpurun_out | grep "${T}_" >&2
