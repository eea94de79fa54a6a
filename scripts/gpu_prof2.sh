#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -c 260 -o gpurun_out/prof_all -f \
    python scripts/profile_iter.py --msteps 2 > gpurun_out/prof_all.log 2>&1
echo "ncu all exit $?"; tail -2 gpurun_out/prof_all.log
SWEEP_STEPS=10 bash scripts/gpu_sweep.sh "" "--batch 12" "--batch 48" "--batch 148 --msteps 8"
