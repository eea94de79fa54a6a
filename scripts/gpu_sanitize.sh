#!/bin/bash
# compute-sanitizer memcheck over one small unrolled iteration of every scene / model (B200, a few minutes)
set -u
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck.log \
    python -m pytest tests/test_gpu_unroll.py tests/test_gpu_burgers.py -m gpu -q -x \
    -k "(64x32m2 and (defaults or simt or stack-solverio)) or mercury or C1-32x32 or rollout" > gpurun_out/memcheck_pytest.log 2>&1
echo "memcheck exit $?"; tail -5 gpurun_out/memcheck_pytest.log; grep -E "ERROR SUMMARY|Invalid|Error" gpurun_out/memcheck.log | head -20
if [ "${RACECHECK:-0}" = "1" ]; then
timeout 500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 --log-file gpurun_out/racecheck.log \
    python -m pytest tests/test_gpu_unroll.py tests/test_gpu_burgers.py -m gpu -q -x -k "(64x32m2 and simt) or C1-32x32" > gpurun_out/racecheck_pytest.log 2>&1
echo "racecheck exit $?"; tail -3 gpurun_out/racecheck_pytest.log; grep -E "RACECHECK SUMMARY|hazard" gpurun_out/racecheck.log | sort | uniq -c | head -20
fi
