#!/bin/bash
# Bundled GPU session: parity suite (short gate first), conv phase timeline, bench with and without PDL.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stages.py -m gpu -q -x > gpurun_out/pytest_gate.log 2>&1
rc=$?; echo "gate exit $rc"; tail -3 gpurun_out/pytest_gate.log
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3
grep -E "^FAILED" gpurun_out/pytest_gpu.log | head -20
echo "== conv phase timeline"; timeout 120 python scripts/tc_trace.py 2>&1 | tail -26
for cfg in "--pdl 1" "--pdl 0" ${EXTRA_CFG:-}; do
echo "== bench $cfg"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $cfg > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; r2=d.get('roofline_pressure_solve') or d.get('roofline_conv') or {}
    print('ms/iter %.2f value %.3e e2e %.3e cg_iters %s launches %d loss %.4f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['config']['mean_cg_iters'],d['gpu_launches'],d['config']['loss']))
    for x in (r,r2):
        if x: print('  roofline %s: %.1f us/launch achieved %.1f %s frac %.4f share %.3f'%(x['kernel'][:24],x['us_per_launch'],x['achieved'],x['unit'],x['frac'],x['share_of_step']))
except Exception as e: print('bench failed', e)
"
done
