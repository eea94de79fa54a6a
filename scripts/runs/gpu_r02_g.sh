#!/bin/bash
mkdir -p gpurun_out/r02_g
O=gpurun_out/r02_g
python scripts/direct_trace.py > $O/direct_trace.txt 2>&1; cat $O/direct_trace.txt
b() { name=$1; shift; python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > $O/bench_$name.json 2> $O/bench_$name.err; python -c "import json;d=json.load(open('$O/bench_$name.json'));print('$name ms_per_step %.3f'%(d['ms_per_step']))" || tail -3 $O/bench_$name.err; }
b bg0 --wgrad-bg-ctas 0
b bg56c16 --wgrad-bg-ctas 56 --wgrad-bg-chunk 16
b bg64c16 --wgrad-bg-ctas 64 --wgrad-bg-chunk 16
b bg40c16 --wgrad-bg-ctas 40 --wgrad-bg-chunk 16
b bg48c12 --wgrad-bg-ctas 48 --wgrad-bg-chunk 12
b bg72c8 --wgrad-bg-ctas 72 --wgrad-bg-chunk 8
timeout 600 python -m pytest tests/test_trainer_options.py -q -m gpu -s 2>&1 | tail -5
