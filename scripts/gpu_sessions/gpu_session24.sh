#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_detconv_gpu.py tests/test_nets_gpu.py -m gpu -q -x -k "detconv or detector" > gpurun_out/pytest_rs.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|Error|assert|rc=" gpurun_out/pytest_rs.log | tail -8
PV_RS_DEBUG=1 timeout 300 python scripts/gpu_probe_det.py --frames 8 --reps 2 2>&1 | tail -1
timeout 300 python scripts/gpu_probe_det.py --frames 8 2>&1 | tail -1
