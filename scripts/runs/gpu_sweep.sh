#!/bin/bash
# bench sweep over option sets given as arguments (one quoted string each)
set -u
mkdir -p gpurun_out
for cfg in "$@"; do
echo "== bench $cfg"
timeout 300 python bench.py --steps ${SWEEP_STEPS:-10} --warmup 3 --no-cpu-baseline $cfg > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('ms/iter %.2f value %.3e e2e %.3e launches %d loss %.4f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['gpu_launches'],d['config']['loss']))
" || tail -5 gpurun_out/bench.log
done
