#!/bin/bash
mkdir -p gpurun_out/r02_i
O=gpurun_out/r02_i
timeout 2400 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
python scripts/direct_trace.py > $O/direct_trace.txt 2>&1; cat $O/direct_trace.txt
b() { name=$1; shift; python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > $O/bench_$name.json 2> $O/bench_$name.err; python -c "import json;d=json.load(open('$O/bench_$name.json'));print('$name ms_per_step %.3f  solve %.2f us'%(d['ms_per_step'],(d.get('roofline_pressure_solve') or d['roofline'])['us_per_launch']))" || tail -3 $O/bench_$name.err; }
b default
b c4 --config c4
