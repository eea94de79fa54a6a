#!/bin/bash
# full-set metrics (raw CSV) of selected kernels of one unrolled step: REGEX / COUNT from the environment
set -u
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --clock-control none -c ${COUNT:-8} \
    -k regex:"${REGEX:-k_direct|k_wgrad_c32_tc}" --csv --page raw --log-file gpurun_out/sel_raw.csv \
    python scripts/profile_iter.py --msteps 2 > gpurun_out/prof_sel.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/prof_sel.log; ls -la gpurun_out/sel_raw.csv
