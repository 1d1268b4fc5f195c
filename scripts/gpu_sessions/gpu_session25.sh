#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nets_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -k "detector" > gpurun_out/pytest_c1.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|Error|assert|rc=" gpurun_out/pytest_c1.log | tail -8
timeout 300 python scripts/gpu_probe_det.py --frames 8 2>&1 | tail -1
