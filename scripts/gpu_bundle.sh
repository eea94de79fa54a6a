#!/bin/bash
# One bundled GPU session: parity tests, benches in several configurations, micro-benchmarks, ncu captures.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3
grep -E "^FAILED" gpurun_out/pytest_gpu.log | head -20
for cfg in "--conv-path 2 --wgrad-path 2 --cg-precond 1" "--conv-path 2 --wgrad-path 1 --cg-precond 0 --cg-rows 16" "--conv-path 2 --wgrad-path 2 --cg-precond 0 --cg-rows 16"; do
  echo "== bench $cfg"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $cfg 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('ms/iter %.2f value %.3e e2e %.3e cg_iters %s roofline_us %.1f launches %d'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['config']['mean_cg_iters'],d['roofline']['us_per_launch'],d['gpu_launches']))
except Exception as e: print('bench failed', e)
"
done
echo "== tc conv timing"; timeout 300 python scripts/tc_debug.py 2>&1 | tail -4
echo "== cg bench"; timeout 300 python scripts/cg_bench.py 2>&1 | grep "128x64" | head -12
echo "== launch list (msteps 8, best config)"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tc.csv python scripts/profile_iter.py --msteps 8 --conv-path 2 --wgrad-path 2 > gpurun_out/prof_launches.log 2>&1; echo "launch-list exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'k_cg_mg|k_conv5x5_c32_tc|k_wgrad_c32_tc' -c 5 -o gpurun_out/prof_top2 -f python scripts/profile_iter.py --msteps 2 --conv-path 2 --wgrad-path 2 > gpurun_out/prof_full.log 2>&1; echo "full exit $?"
ls gpurun_out | head -30
