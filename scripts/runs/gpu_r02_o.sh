#!/bin/bash
OUT=gpurun_out/r02_o; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_stages.py -x -q -m gpu -k "conv5x5" > $OUT/pytest_conv.log 2>&1; tail -5 $OUT/pytest_conv.log
timeout 900 python -m pytest tests/test_gpu_unroll.py tests/test_gpu_quoted_configs.py tests/test_gpu_burgers.py -x -q -m gpu > $OUT/pytest_unroll.log 2>&1; tail -5 $OUT/pytest_unroll.log
timeout 300 python scripts/chain_trace.py > $OUT/chain_trace_sol32.txt 2> $OUT/chain_trace.err; cat $OUT/chain_trace_sol32.txt; tail -3 $OUT/chain_trace.err
timeout 300 python scripts/chain_trace.py --opt thin_path=1 --opt pdl=0 > $OUT/chain_trace_sol32_oldthin_nopdl.txt 2>> $OUT/chain_trace.err; cat $OUT/chain_trace_sol32_oldthin_nopdl.txt
python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; python -c "import json;d=json.load(open('$OUT/bench.json'));print('ms_per_step %.3f e2e %.3f'%(d['ms_per_step'],d['e2e']['ms_per_step']))"
