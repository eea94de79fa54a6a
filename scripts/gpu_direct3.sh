#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k 'pressure_solve or reference_style or step_forward' 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_direct -s 12 -c 4 python scripts/direct_bench.py 2>&1 | grep -E "k_direct|gpu__time|issue_active" | head -16
SWEEP_STEPS=20 bash scripts/gpu_sweep.sh "--direct-solve 1" "--direct-solve 0"
