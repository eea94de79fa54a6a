#!/bin/bash
mkdir -p gpurun_out/r02_f
O=gpurun_out/r02_f
timeout 2400 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
grep "weight gradient\|c4 step  3" $O/pytest_gpu.log | cut -c1-200
b() { name=$1; shift; python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > $O/bench_$name.json 2> $O/bench_$name.err; python -c "import json;d=json.load(open('$O/bench_$name.json'));print('$name ms_per_step %.3f  conv %.2f us  solve %.2f us  loss %.5f'%(d['ms_per_step'],d['roofline']['us_per_launch'] if 'conv' in d['roofline']['kernel'] else -1,(d.get('roofline_pressure_solve') or d['roofline'])['us_per_launch'], d['config']['loss']))" || tail -3 $O/bench_$name.err; }
b default
b bg0 --wgrad-bg-ctas 0
b bg32 --wgrad-bg-ctas 32
b bg52 --wgrad-bg-ctas 52
b bg48c2 --wgrad-bg-ctas 48 --wgrad-bg-chunk 2
b bg48c8 --wgrad-bg-ctas 48 --wgrad-bg-chunk 8
b c4 --config c4
b c4_iter --config c4 --direct-solve 0
b c2 --config c2
