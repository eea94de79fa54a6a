#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "pressure_solve or reference_style or step_forward" > gpurun_out/pytest_new.log 2>&1
echo "pytest stages exit $?"; tail -8 gpurun_out/pytest_new.log
timeout 600 python -m pytest tests/test_gpu_unroll.py -m gpu -q -x -k "defaults or simt or full_size or graph or rollout" > gpurun_out/pytest_new2.log 2>&1
echo "pytest unroll exit $?"; tail -8 gpurun_out/pytest_new2.log
SWEEP_STEPS=20 bash scripts/gpu_sweep.sh "--direct-solve 0" "--direct-solve 1" "--direct-solve 1 --wgrad-overlap 0"
