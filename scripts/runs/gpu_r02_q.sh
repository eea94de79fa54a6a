#!/bin/bash
OUT=gpurun_out/r02_q; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_unroll.py -x -q -m gpu -s -k "packed or graph or full_size" > $OUT/pytest_packed.log 2>&1; grep -E "iteration|passed|failed|Error|assert" $OUT/pytest_packed.log | head -30
timeout 900 python -m pytest tests/test_gpu_quoted_configs.py tests/test_gpu_stages.py -x -q -m gpu > $OUT/pytest_quoted.log 2>&1; tail -4 $OUT/pytest_quoted.log
timeout 300 python scripts/chain_trace.py > $OUT/chain_trace_sol32.txt 2> $OUT/chain_trace.err; head -12 $OUT/chain_trace_sol32.txt; tail -3 $OUT/chain_trace.err
timeout 300 python scripts/chain_trace.py --opt act_pack=0 > $OUT/chain_trace_sol32_nopack.txt 2>> $OUT/chain_trace.err; head -5 $OUT/chain_trace_sol32_nopack.txt
python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; python -c "import json;d=json.load(open('$OUT/bench.json'));print('ms_per_step %.3f e2e %.3f conv %.2f us'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['us_per_launch']))"
