#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_stages.py tests/test_gpu_unroll.py -m gpu -q -x > gpurun_out/pytest_gate.log 2>&1
echo "gate exit $?"; tail -3 gpurun_out/pytest_gate.log; grep -E "^FAILED|Error" gpurun_out/pytest_gate.log | head
echo "== cg bench"; timeout 300 python scripts/cg_bench.py 2>&1 | grep -E "128x64 B=  3 cluster=1 rows=16|128x64 B=148 cluster=1 rows=16" | head -8
bash scripts/gpu_sweep.sh "--wgrad-overlap 0" "--wgrad-window-us 110" "--wgrad-window-us 160" "--wgrad-window-us 220"
