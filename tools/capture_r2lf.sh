#!/usr/bin/env bash
set -u
T=${1:-r2lf}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_longform.py tests/test_gpu_ops.py tests/test_gpu_engine.py -q --durations=5 2>&1 | tail -40 | tee gpurun_out/${T}_tests.log
