#!/usr/bin/env bash
# One-call evidence capture for a round (run under gpurun from the repo root):
#   gpurun --timeout 1800 -- 'bash tools/capture_round.sh r2z'
# Writes small text files under gpurun_out/<tag>_* (ncu reports stay in /tmp on the box: gpurun_out is capped at 64 MiB);
# copy what should be judged into profiles/.  Order: parity first, then the contract bench (its extras carry configs
# 3 / 4 / 5), then the profiler passes (numbers printed by a run under ncu are never bench values).
set -u
T=${1:-rX}
mkdir -p gpurun_out
step() { echo "== $1" >&2; }
step "gpu test suite";        timeout 900 python -m pytest tests -m gpu -x -q --durations=10 -s > gpurun_out/${T}_gpu_suite.log 2>&1; tail -3 gpurun_out/${T}_gpu_suite.log
step "smoke";                 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/${T}_smoke.log
step "bench N=1";             timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench.err; echo "rc=$?"; cut -c1-300 gpurun_out/${T}_bench_1gpu.json
step "reference arm";         timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; cut -c1-260 gpurun_out/${T}_bench_reference.json
step "kernel micro-benches";  timeout 150 python tools/bench_kernels.py 24 2>/dev/null | grep '^{' > gpurun_out/${T}_kernels.jsonl; cut -c1-200 gpurun_out/${T}_kernels.jsonl
step "beams / word mode";     timeout 300 python tools/bench_beams.py 2>/dev/null | grep '^{' > gpurun_out/${T}_beams.jsonl; cat gpurun_out/${T}_beams.jsonl; timeout 200 python tools/bench_word_mode.py 4 2>/dev/null | grep '^{' > gpurun_out/${T}_word_mode_e2e.jsonl; cat gpurun_out/${T}_word_mode_e2e.jsonl
step "ncu launch list";       timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches_ncu.csv > gpurun_out/${T}_launches_summary.md 2>/dev/null
step "ncu --set full"
timeout 400 ncu --set full --clock-control none -k regex:"decode_attn|lmhead|skinny|finalize|embed" -s 35 -c 36 -o /tmp/${T}_decode -f python tools/ncu_target.py 24 3 0 > gpurun_out/${T}_ncu_decode.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"gemm_bf16|attention_enc|layernorm_kernel|logmel" -c 12 -o /tmp/${T}_encoder -f python tools/ncu_target.py 24 0 1 > gpurun_out/${T}_ncu_encoder.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__waves_per_multiprocessor,lts__t_bytes.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,sm__cycles_elapsed.avg.per_second
for r in decode encoder; do
  ncu -i /tmp/${T}_$r.ncu-rep --page raw --csv --metrics $M > gpurun_out/${T}_ncu_${r}_raw.csv 2>/dev/null
done
python tools/ncu_traffic.py gpurun_out/${T}_ncu_decode_raw.csv gpurun_out/${T}_ncu_encoder_raw.csv > gpurun_out/${T}_ncu_traffic.json 2>gpurun_out/${T}_ncu_traffic.err
step "library baseline (transformers on the same GPU)"; timeout 300 python tools/hf_gpu_baseline.py 24 2>/dev/null | grep '^{' > gpurun_out/${T}_hf_gpu_baseline.json; cat gpurun_out/${T}_hf_gpu_baseline.json
du -sh gpurun_out >&2
