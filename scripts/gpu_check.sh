#!/bin/bash
# Runs on the B200 box via gpurun: GPU parity tests, smoke, a short bench; logs into gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3
if [ "${SKIP_SMOKE:-0}" != "1" ]; then
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log
if [ -n "${EXTRA:-}" ]; then
timeout 900 bash -c "$EXTRA" > gpurun_out/extra.log 2>&1
echo "extra exit $?" >> gpurun_out/extra.log
tail -60 gpurun_out/extra.log
fi
