#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stages.py tests/test_gpu_unroll.py -m gpu -q -x > gpurun_out/pytest_gate.log 2>&1
echo "gate exit $?"; tail -2 gpurun_out/pytest_gate.log
for cfg in "--pdl 1 --graph 1" "--pdl 0 --graph 1" "--pdl 1 --graph 0"; do
  echo "== graph trace $cfg"; timeout 200 python scripts/graph_trace.py $cfg 2>&1 | tail -12
done
echo "== bench"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; r2=d.get('roofline_pressure_solve') or d.get('roofline_conv') or {}
print('ms/iter %.2f value %.3e e2e %.3e launches %d loss %.4f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['gpu_launches'],d['config']['loss']))
for x in (r,r2):
    if x: print('  roofline %s: %.1f us/launch achieved %.1f %s frac %.4f share %.3f'%(x['kernel'][:24],x['us_per_launch'],x['achieved'],x['unit'],x['frac'],x['share_of_step']))
"
