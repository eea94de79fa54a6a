#!/bin/bash
# r02_m: tightened parity asserts + graph re-capture test, in-chain timeline of one SOL-32 iteration, lanes re-measured
OUT=gpurun_out/r02_m; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_unroll.py tests/test_gpu_stages.py tests/test_gpu_compat.py tests/test_gpu_burgers.py tests/test_gpu_quoted_configs.py -x -q -m gpu > $OUT/pytest_sel.log 2>&1; tail -15 $OUT/pytest_sel.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 300 python scripts/chain_trace.py > $OUT/chain_trace_sol32.txt 2> $OUT/chain_trace.err; cat $OUT/chain_trace_sol32.txt; tail -3 $OUT/chain_trace.err
timeout 300 python scripts/probes/lane_probe2.py > $OUT/lanes.txt 2>&1; cat $OUT/lanes.txt
