#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 120 python scripts/stack_probe.py 2>&1 | grep -E "capacity|stack [01]" 
echo "probe exit $?"
timeout 600 python -m pytest tests/test_gpu_unroll.py -m gpu -q -x > gpurun_out/pytest_new.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_new.log
SWEEP_STEPS=20 bash scripts/gpu_sweep.sh "--conv-stack 0" "--conv-stack 1" "--conv-stack 0" "--conv-stack 1"
