#!/bin/bash
OUT=gpurun_out/r02_y; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_compat.py tests/test_gpu_pipelines.py -q -m gpu -s -k "any_grid or plugin or trajectory" > $OUT/pytest_new.log 2>&1; grep -E "step rel|step bwd|unroll losses|generated trajectory|passed|failed|Error|assert " $OUT/pytest_new.log | head -40
