#!/usr/bin/env bash
# r2b: bench line + ncu evidence (small) + attention variants + MUFU probe
set -u
T=${1:-r2b}
mkdir -p gpurun_out
step() { echo "== $1" >&2; }
step "bench N=1";  timeout 700 python bench.py --steps 8 --warmup 3 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench.err; echo "rc=$?"; cut -c1-300 gpurun_out/${T}_bench_1gpu.json
step "attention variants"
for v in default max3 poly4 poly4max3 poly3max3; do
  if [ "$v" = default ]; then unset TWB200_LIB; else export TWB200_LIB=$PWD/turbo-whisper-workspace_b200/variants/libtwb200_$v.so; fi
  timeout 120 python tools/bench_kernels.py 24 attention 2>/dev/null | grep '^{' | sed "s/^{/{\"variant\": \"$v\", /" >> gpurun_out/${T}_attn_variants.jsonl
done
unset TWB200_LIB
cat gpurun_out/${T}_attn_variants.jsonl
step "MUFU probe"; (cd tools/probes && nvcc -O3 -gencode arch=compute_100a,code=sm_100a ex2_packed_probe.cu -o /tmp/ex2p && /tmp/ex2p) > gpurun_out/${T}_mufu_probe.txt 2>&1; cat gpurun_out/${T}_mufu_probe.txt
step "ncu launch list"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches_ncu.csv > gpurun_out/${T}_launches_summary.md 2>/dev/null
step "ncu --set full"
timeout 300 ncu --set full --clock-control none -k regex:"decode_attn|lmhead|skinny|finalize|embed" -c 22 -o /tmp/${T}_decode -f python tools/ncu_target.py 24 2 0 > gpurun_out/${T}_ncu_decode.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"gemm_bf16|attention_enc|layernorm_kernel|logmel" -c 12 -o /tmp/${T}_encoder -f python tools/ncu_target.py 24 0 1 > gpurun_out/${T}_ncu_encoder.log 2>&1
python tools/ncu_traffic.py /tmp/${T}_decode.ncu-rep /tmp/${T}_encoder.ncu-rep > gpurun_out/${T}_ncu_traffic.json 2>gpurun_out/${T}_ncu_traffic.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,lts__t_bytes.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,sm__cycles_elapsed.avg.per_second
for r in decode encoder; do
  ncu -i /tmp/${T}_$r.ncu-rep --page raw --csv --metrics $M > gpurun_out/${T}_ncu_${r}_raw.csv 2>/dev/null
done
ls -la /tmp/*.ncu-rep >&2
du -sh gpurun_out >&2
