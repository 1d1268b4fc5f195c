#!/bin/bash
# c12 bring-up: unit tests, then per-layer probe in both modes (+ role counters)
TAG=${1:-c12a}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_c12_gpu.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -30 gpurun_out/${TAG}_pytest.log
PV_DET_CONV1=c12 PV_C12_DEBUG=1 timeout 300 python scripts/gpu_probe_det.py --frames 8 > gpurun_out/${TAG}_probe_c12.json 2>&1
tail -n 2 gpurun_out/${TAG}_probe_c12.json
