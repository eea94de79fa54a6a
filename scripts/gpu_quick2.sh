#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_burgers.py tests/test_gpu_stages.py -m gpu -q -x -k "burgers or 256x128" > gpurun_out/pytest_new.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_new.log
timeout 300 python scripts/lane_probe.py > gpurun_out/lane_probe.log 2>&1
echo "probe exit $?"; cat gpurun_out/lane_probe.log
