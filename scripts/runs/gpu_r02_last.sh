#!/bin/bash
OUT=gpurun_out/r02_last; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 --log-file $OUT/racecheck.log \
    python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "thin_paths or any_grid_size and 40x20" > $OUT/racecheck_pytest.log 2>&1
echo "racecheck exit $?"; tail -3 $OUT/racecheck_pytest.log; grep -E "RACECHECK SUMMARY|hazard" $OUT/racecheck.log | sort | uniq -c | head -10
timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
