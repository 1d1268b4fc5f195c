#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -x -k "c2 or c4" > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|Error|assert" gpurun_out/pytest_f.log | tail -5
