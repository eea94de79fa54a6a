#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "mercury or phi2 or conv5x5 or burgers" > gpurun_out/pytest_new.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/pytest_new.log
