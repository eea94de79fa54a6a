#!/bin/bash
# One bundled GPU session: parity tests, benches, micro-benchmarks, ncu launch list.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3
grep -E "^FAILED" gpurun_out/pytest_gpu.log | head -20
for cfg in "" "--cg-rows 8" "--cg-precond 0 --cg-rows 16"; do
  echo "== bench $cfg"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $cfg 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('ms/iter %.2f value %.3e e2e %.3e cg_iters %s roofline_us %.1f launches %d loss %.4f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['config']['mean_cg_iters'],d['roofline']['us_per_launch'],d['gpu_launches'],d['config']['loss']))
except Exception as e: print('bench failed', e)
"
done
echo "== cg bench"; timeout 300 python scripts/cg_bench.py 2>&1 | grep "128x64 B=  3" | head -8
echo "== launch list (msteps 8)"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tc.csv python scripts/profile_iter.py --msteps 8 > gpurun_out/prof_launches.log 2>&1; echo "launch-list exit $?"
python - <<'PY'
import csv, collections, re
lines=[l for l in open('gpurun_out/launches_tc.csv') if not l.startswith('==')]
agg=collections.defaultdict(lambda:[0,0.0]); tot=0
for row in csv.DictReader(lines):
    name=re.sub(r'\(.*','',row['Kernel Name']); v=float(row['Metric Value'].replace(',',''))
    if row['Metric Unit']=='ns': v/=1000.0
    agg[name][0]+=1; agg[name][1]+=v; tot+=v
print("total us %.1f launches %d"%(tot,sum(a[0] for a in agg.values())))
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]:
    print("%-44s n=%4d total=%9.1f avg=%8.2f share=%5.1f%%"%(k.replace('void ','').strip()[:44],n,t,t/n,100*t/tot))
PY
