#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_stages.py tests/test_gpu_unroll.py -m gpu -q -x > gpurun_out/pytest_gate.log 2>&1
echo "gate exit $?"; tail -3 gpurun_out/pytest_gate.log; grep -E "^FAILED|Error" gpurun_out/pytest_gate.log | head
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -2; grep -E "^FAILED" gpurun_out/pytest_gpu.log | head
for cfg in "--pdl 1 --chain 0" "--pdl 1 --chain 1" "--pdl 0 --chain 0"; do
  echo "== graph trace $cfg"; timeout 200 python scripts/graph_trace.py $cfg 2>&1 | tail -13
done
for cfg in "--conv-chain 0" "--conv-chain 1"; do
echo "== bench $cfg"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $cfg > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; r2=d.get('roofline_pressure_solve') or d.get('roofline_conv') or {}
print('ms/iter %.2f value %.3e e2e %.3e launches %d loss %.4f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['gpu_launches'],d['config']['loss']))
for x in (r,r2):
    if x: print('  roofline %s: %.1f us/launch achieved %.1f %s frac %.4f share %.3f'%(x['kernel'][:24],x['us_per_launch'],x['achieved'],x['unit'],x['frac'],x['share_of_step']))
"
done
echo "== launch list (msteps 8)"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01d.csv python scripts/profile_iter.py --msteps 8 > gpurun_out/prof_launches.log 2>&1; echo "launch-list exit $?"
python scripts/launch_summary.py gpurun_out/launches_r01d.csv 12
