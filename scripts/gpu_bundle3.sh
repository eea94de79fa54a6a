#!/bin/bash
# Bundled GPU session: full parity suite, conv phase timeline, bench, ncu captures of the tensor-core kernels, launch list.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3
grep -E "^FAILED" gpurun_out/pytest_gpu.log | head -20
echo "== conv phase timeline"; timeout 120 python scripts/tc_trace.py 2>&1 | tail -30
echo "== bench"
timeout 300 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; r2=d.get('roofline_pressure_solve') or d.get('roofline_conv') or {}
    print('ms/iter %.2f value %.3e e2e %.3e cg_iters %s launches %d loss %.4f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['config']['mean_cg_iters'],d['gpu_launches'],d['config']['loss']))
    for x in (r,r2):
        if x: print('  roofline %s: %.1f us/launch achieved %.1f %s frac %.4f share %.3f'%(x['kernel'][:24],x['us_per_launch'],x['achieved'],x['unit'],x['frac'],x['share_of_step']))
    print('  cpu_baseline', d.get('cpu_baseline',{}).get('value'))
except Exception as e: print('bench failed', e)
"
echo "== ncu full: conv tc + wgrad tc (msteps 2 iteration)"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_conv5x5_c32_tc|k_wgrad_c32_tc|k_conv5x5_expand|k_conv5x5_reduce" -c 30 -o gpurun_out/prof_conv -f python scripts/profile_iter.py --msteps 2 > gpurun_out/prof_conv.log 2>&1; echo "ncu exit $?"
echo "== launch list (msteps 8)"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01c.csv python scripts/profile_iter.py --msteps 8 > gpurun_out/prof_launches.log 2>&1; echo "launch-list exit $?"
python scripts/launch_summary.py gpurun_out/launches_r01c.csv 24
