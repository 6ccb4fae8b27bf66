#!/usr/bin/env bash
set -u
T=${1:-r2final}
mkdir -p gpurun_out
timeout 230 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${T}_gpu_suite_tail.log
