#!/bin/bash
OUT=gpurun_out/r02_p; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_unroll.py tests/test_gpu_quoted_configs.py tests/test_gpu_burgers.py tests/test_gpu_compat.py tests/test_gpu_errors.py tests/test_gpu_pipelines.py -q -m gpu > $OUT/pytest_unroll.log 2>&1; tail -8 $OUT/pytest_unroll.log
timeout 300 python scripts/chain_trace.py > $OUT/chain_trace_sol32.txt 2> $OUT/chain_trace.err; head -14 $OUT/chain_trace_sol32.txt; tail -3 $OUT/chain_trace.err
timeout 300 python scripts/chain_trace.py --opt thin_late_trigger=1 > $OUT/chain_trace_sol32_late.txt 2>> $OUT/chain_trace.err; head -14 $OUT/chain_trace_sol32_late.txt
