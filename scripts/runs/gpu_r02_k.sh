#!/bin/bash
mkdir -p gpurun_out/r02_k
O=gpurun_out/r02_k
timeout 900 python -m pytest tests/test_gpu_unroll.py tests/test_gpu_stages.py tests/test_gpu_quoted_configs.py -q -m gpu -x > $O/pytest_sel.log 2>&1; tail -3 $O/pytest_sel.log
grep "weight gradient" $O/pytest_sel.log | cut -c1-200
b() { name=$1; shift; python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > $O/bench_$name.json 2> $O/bench_$name.err; python -c "import json;d=json.load(open('$O/bench_$name.json'));print('$name ms_per_step %.3f'%(d['ms_per_step']))" || tail -3 $O/bench_$name.err; }
b issuers2
SOL_WGRAD_ISSUERS=1 b issuers1
