#!/bin/bash
mkdir -p gpurun_out/r02_h
O=gpurun_out/r02_h
timeout 2400 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
grep "weight gradient\|c4 step  3" $O/pytest_gpu.log | cut -c1-200
python scripts/direct_trace.py > $O/direct_trace.txt 2>&1; cat $O/direct_trace.txt
b() { name=$1; shift; python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > $O/bench_$name.json 2> $O/bench_$name.err; python -c "import json;d=json.load(open('$O/bench_$name.json'));print('$name ms_per_step %.3f  solve %.2f us'%(d['ms_per_step'],(d.get('roofline_pressure_solve') or d['roofline'])['us_per_launch']))" || tail -3 $O/bench_$name.err; }
b default
b nofuse --fuse-stencil 0
b c4 --config c4
