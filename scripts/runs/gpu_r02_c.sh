#!/bin/bash
set -x
mkdir -p gpurun_out
for v in 2 3 4 5 6; do
  python scripts/conv_trace.py $v > gpurun_out/r02_c_trace_v$v.txt 2>&1
done
cat gpurun_out/r02_c_trace_v2.txt
for v in 3 4 5 6; do head -3 gpurun_out/r02_c_trace_v$v.txt; done
python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
from solver_in_the_loop_b200 import engine
import subprocess
PY
for v in 2 3 4 5 6; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --conv-variant $v > gpurun_out/r02_c_bench_v$v.json 2> gpurun_out/r02_c_bench_v$v.err
  python -c "import json;d=json.load(open('gpurun_out/r02_c_bench_v$v.json'));print('variant $v ms_per_step',d['ms_per_step'],'conv us',d['roofline']['us_per_launch'])"
done
