#!/bin/bash
# round-2 bundle B: new 3xFP16 conv kernel — numerics, variants timing, unrolled parity, bench
set -x
mkdir -p gpurun_out
python scripts/conv_bench.py > gpurun_out/r02_b_conv_bench.txt 2>&1
cat gpurun_out/r02_b_conv_bench.txt
timeout 900 python -m pytest tests/test_gpu_stages.py -q -x -m gpu -k "conv5x5" -s > gpurun_out/r02_b_stages_conv.log 2>&1
tail -5 gpurun_out/r02_b_stages_conv.log
timeout 1500 python -m pytest tests/test_gpu_unroll.py tests/test_gpu_quoted_configs.py -q -m gpu -s > gpurun_out/r02_b_unroll.log 2>&1
tail -5 gpurun_out/r02_b_unroll.log
for v in 0 1 2; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --conv-variant $v > gpurun_out/r02_b_bench_v$v.json 2> gpurun_out/r02_b_bench_v$v.err
  python -c "import json;d=json.load(open('gpurun_out/r02_b_bench_v$v.json'));print('variant $v ms_per_step',d['ms_per_step'],'conv us',d['roofline']['us_per_launch'])"
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --conv-path 3 > gpurun_out/r02_b_bench_tf32.json 2> gpurun_out/r02_b_bench_tf32.err
python -c "import json;d=json.load(open('gpurun_out/r02_b_bench_tf32.json'));print('tf32 ms_per_step',d['ms_per_step'],'conv us',d['roofline']['us_per_launch'])"
