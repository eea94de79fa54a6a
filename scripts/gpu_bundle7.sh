#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gap probe"; timeout 120 ./scripts/probes/gap_probe 2>&1 | grep -E "UMMA|TMEM alloc" 
timeout 500 python -m pytest tests/test_gpu_stages.py tests/test_gpu_unroll.py -m gpu -q -x > gpurun_out/pytest_gate.log 2>&1
echo "gate exit $?"; tail -3 gpurun_out/pytest_gate.log; grep -E "^FAILED|Error" gpurun_out/pytest_gate.log | head
for cfg in "--wgrad-overlap 1" "--wgrad-overlap 0"; do
echo "== bench $cfg"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $cfg > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; r2=d.get('roofline_pressure_solve') or d.get('roofline_conv') or {}
print('ms/iter %.2f value %.3e e2e %.3e launches %d loss %.4f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['gpu_launches'],d['config']['loss']))
" || tail -5 gpurun_out/bench.log
done
