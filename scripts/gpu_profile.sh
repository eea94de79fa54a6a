#!/bin/bash
# ncu launch list of one training iteration + full-set captures of the top kernels (1 GPU).
set -u
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python scripts/profile_iter.py --msteps 8 > gpurun_out/prof_launches.log 2>&1
echo "launch-list exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'k_cg|k_conv5x5_c32|k_wgrad_c32' -c 6 -o gpurun_out/prof_top -f python scripts/profile_iter.py --msteps 2 > gpurun_out/prof_full.log 2>&1
echo "full exit $?"
ls -la gpurun_out
