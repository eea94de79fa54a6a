#!/bin/bash
# full-set metrics of every kernel kind EXCEPT the two already captured (conv tc, cg mg3), one unrolled step, as raw CSV
set -u
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --clock-control none -c 48 \
    -k regex:'^(?!.*(k_conv5x5_c32_tc|k_cg_mg3|k_prep_tc|k_flip|k_wgrad_finalize)).*$' --csv --page raw --log-file gpurun_out/all_raw.csv \
    python scripts/profile_iter.py --msteps 1 > gpurun_out/prof_all.log 2>&1
echo "ncu all exit $?"; tail -2 gpurun_out/prof_all.log; ls -la gpurun_out; head -c 600 gpurun_out/all_raw.csv
