#!/usr/bin/env bash
# One-call evidence capture for a round (run under gpurun from the repo root):
#   gpurun --timeout 900 -- 'bash tools/capture_round.sh r2a'
# Writes everything under gpurun_out/<tag>_*; copy what should be judged into profiles/.
# Order: parity first, then the contract bench, then the other BASELINE.json configs, then profiler passes
# (numbers printed by a run under ncu are never bench values).
set -u
T=${1:-rX}
mkdir -p gpurun_out
step() { echo "== $1" >&2; }

step "gpu test suite";        timeout 300 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${T}_gpu_suite.log 2>&1; tail -2 gpurun_out/${T}_gpu_suite.log
step "bench N=1";             timeout 240 python bench.py --steps 8 --warmup 3 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench.err; cut -c1-200 gpurun_out/${T}_bench_1gpu.json
step "config 3 (1 h file)";   timeout 200 python tools/run_configs.py 3 2>/dev/null | grep '^{' > gpurun_out/${T}_config3.jsonl; cat gpurun_out/${T}_config3.jsonl
step "config 4 (large-v3)";   timeout 240 python tools/run_configs.py 4 2>/dev/null | grep '^{' > gpurun_out/${T}_config4.jsonl; cat gpurun_out/${T}_config4.jsonl
step "word-mode e2e";         timeout 150 python tools/bench_word_mode.py 4 2>/dev/null | grep '^{' > gpurun_out/${T}_word_mode_e2e.jsonl; cat gpurun_out/${T}_word_mode_e2e.jsonl
step "kernel micro-benches";  timeout 150 python tools/bench_kernels.py 24 2>/dev/null | grep '^{' > gpurun_out/${T}_kernels.jsonl; timeout 60 python tools/bench_word_align.py 24 2>/dev/null | grep '^{' >> gpurun_out/${T}_kernels.jsonl
step "ncu launch list";       timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
[ -f tools/summarize_launches.py ] && python tools/summarize_launches.py gpurun_out/${T}_launches_ncu.csv > gpurun_out/${T}_launches_summary.md 2>/dev/null
ls -la gpurun_out | grep "${T}_" >&2
