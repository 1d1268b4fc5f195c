#!/bin/bash
# detconv bring-up: unit parity, detector parity, per-layer probe for both conv implementations
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_detconv_gpu.py -m gpu -q -x > gpurun_out/pytest_detconv.log 2>&1; echo "detconv pytest rc=$?"
grep -E "passed|failed|Error|assert|rc=" gpurun_out/pytest_detconv.log | tail -12
timeout 600 python -m pytest tests/test_nets_gpu.py -m gpu -q -x -k "detector" > gpurun_out/pytest_det.log 2>&1; echo "detector pytest rc=$?"
grep -E "passed|failed|Error|assert" gpurun_out/pytest_det.log | tail -8
timeout 300 python scripts/gpu_probe_det.py --frames 8 2>&1 | tail -1
PV_DET_CONVS=srgemm timeout 300 python scripts/gpu_probe_det.py --frames 8 2>&1 | tail -1
