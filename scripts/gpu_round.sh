#!/bin/bash
# One-call GPU bundle: parity tests, smoke, both bench arms, ncu launch list (+ optional full-set capture of the top kernels).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log; grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu.log | head -20
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.log 2>&1
echo "bench exit $?" | tee -a gpurun_out/bench.log; tail -2 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
echo "bench ref exit $?" | tee -a gpurun_out/bench_ref.log; tail -2 gpurun_out/bench_ref.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python scripts/profile_iter.py --msteps 32 > gpurun_out/prof_launches.log 2>&1
echo "launch-list exit $?"; tail -2 gpurun_out/prof_launches.log
if [ "${FULL:-0}" = "1" ]; then
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'k_cg_mg3|k_conv5x5_c32_tc|k_wgrad_c32_tc' -c 9 -o gpurun_out/prof_top -f python scripts/profile_iter.py --msteps 2 > gpurun_out/prof_full.log 2>&1
echo "full exit $?"; tail -2 gpurun_out/prof_full.log
fi
if [ -n "${EXTRA:-}" ]; then
timeout ${EXTRA_TIMEOUT:-600} bash -c "$EXTRA" > gpurun_out/extra.log 2>&1
echo "extra exit $?" | tee -a gpurun_out/extra.log; tail -40 gpurun_out/extra.log
fi
ls -la gpurun_out
