#!/bin/bash
OUT=gpurun_out/r02_s; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_stages.py -x -q -m gpu -k "conv5x5" > $OUT/pytest_conv.log 2>&1; tail -3 $OUT/pytest_conv.log
timeout 300 python scripts/chain_trace.py > $OUT/chain_trace_sol32.txt 2> $OUT/chain_trace.err; head -24 $OUT/chain_trace_sol32.txt; tail -3 $OUT/chain_trace.err
python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; python -c "import json;d=json.load(open('$OUT/bench.json'));print('ms_per_step %.3f e2e %.3f conv %.2f us'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['us_per_launch']))"
