#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_unroll.py -m gpu -q -x > gpurun_out/pytest_gate.log 2>&1
echo "gate exit $?"; tail -3 gpurun_out/pytest_gate.log; grep -E "^FAILED|Error" gpurun_out/pytest_gate.log | head
echo "== mg regular probe"; timeout 120 python scripts/mg_regular_probe.py 2>&1 | tail -3
echo "== conv trace"; timeout 120 python scripts/tc_trace.py 2>&1 | tail -22
bash scripts/gpu_sweep.sh "--fuse-small 1" "--fuse-small 0"
