#!/bin/bash
mkdir -p gpurun_out/r02_l
O=gpurun_out/r02_l
timeout 2400 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
b() { name=$1; shift; python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > $O/bench_$name.json 2> $O/bench_$name.err; python -c "import json;d=json.load(open('$O/bench_$name.json'));print('$name ms_per_step %.3f'%(d['ms_per_step']))" || tail -3 $O/bench_$name.err; }
b default
