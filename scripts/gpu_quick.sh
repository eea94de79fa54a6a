#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== wgrad debug"; timeout 300 python scripts/wgrad_debug.py 2>&1 | grep -vE "^   dy" | tail -12
timeout 600 python -m pytest tests/test_gpu_unroll.py tests/test_golden.py -m gpu -q -s > gpurun_out/pytest_unroll.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_unroll.log | tail -2; grep -E "^FAILED|grad rel" gpurun_out/pytest_unroll.log | head -12
for cfg in "--conv-path 2 --wgrad-path 2 --cg-precond 1" "--conv-path 2 --wgrad-path 1 --cg-precond 1"; do
  echo "== bench $cfg"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $cfg 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('ms/iter %.2f value %.3e e2e %.3e cg_iters %s roofline_us %.1f launches %d loss %.4f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['config']['mean_cg_iters'],d['roofline']['us_per_launch'],d['gpu_launches'],d['config']['loss']))
except Exception as e: print('bench failed', e)
"
done
