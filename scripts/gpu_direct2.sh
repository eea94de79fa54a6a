#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:k_direct -s 12 -c 6 python scripts/direct_bench.py 2>&1 | grep -E "k_direct|gpu__time|issue_active|inst_executed|bank_conflicts" | head -40
SWEEP_STEPS=20 bash scripts/gpu_sweep.sh "--direct-solve 1" "--direct-solve 0"
