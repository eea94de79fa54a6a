#!/bin/bash
OUT=gpurun_out/r02_v; mkdir -p $OUT
timeout 300 python scripts/chain_trace.py > $OUT/chain_trace_sol32.txt 2> $OUT/chain_trace.err; head -22 $OUT/chain_trace_sol32.txt; tail -3 $OUT/chain_trace.err
timeout 300 python scripts/chain_trace.py --opt max_carveout=0 > $OUT/chain_trace_sol32_nocarve.txt 2>> $OUT/chain_trace.err; head -8 $OUT/chain_trace_sol32_nocarve.txt
python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; python -c "import json;d=json.load(open('$OUT/bench.json'));print('ms_per_step %.3f e2e %.3f conv %.2f us'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['us_per_launch']))"
