#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py -q -x -m gpu -k "conv5x5" -s > gpurun_out/r02_d_stages_conv.log 2>&1
tail -4 gpurun_out/r02_d_stages_conv.log
grep "wgrad rel" gpurun_out/r02_d_stages_conv.log | head -4
timeout 1500 python -m pytest tests/test_gpu_unroll.py tests/test_gpu_quoted_configs.py -q -m gpu -s > gpurun_out/r02_d_unroll.log 2>&1
tail -4 gpurun_out/r02_d_unroll.log
grep "weight gradient" gpurun_out/r02_d_unroll.log | cut -c1-220
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_d_bench.json 2> gpurun_out/r02_d_bench.err
python -c "import json;d=json.load(open('gpurun_out/r02_d_bench.json'));print('fp16 wgrad ms_per_step',d['ms_per_step'],'conv us',d['roofline']['us_per_launch'], 'solve us', d['roofline_pressure_solve']['us_per_launch'])"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --wgrad-path 3 > gpurun_out/r02_d_bench_w3.json 2> gpurun_out/r02_d_bench_w3.err
python -c "import json;d=json.load(open('gpurun_out/r02_d_bench_w3.json'));print('tf32 wgrad ms_per_step',d['ms_per_step'])"
tail -3 gpurun_out/r02_d_bench.err
