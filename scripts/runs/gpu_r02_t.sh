#!/bin/bash
OUT=gpurun_out/r02_t; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_stages.py -x -q -m gpu > $OUT/pytest_stages.log 2>&1; tail -3 $OUT/pytest_stages.log
for m in 7 3 1; do
timeout 300 python scripts/chain_trace.py --opt thin_cap=$m > $OUT/chain_trace_sol32_cap$m.txt 2> $OUT/chain_trace.err; head -11 $OUT/chain_trace_sol32_cap$m.txt; tail -3 $OUT/chain_trace.err
done
python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; python -c "import json;d=json.load(open('$OUT/bench.json'));print('ms_per_step %.3f e2e %.3f conv %.2f us'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['us_per_launch']))"
