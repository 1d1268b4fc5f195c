#!/bin/bash
TAG=${1:-pyr}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nets_gpu.py tests/test_golden_gpu.py -x -q -k "pyramid or golden or detector_matches" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -8 gpurun_out/${TAG}_pytest.log
PV_RESIZE_PACKED=0 timeout 300 python scripts/gpu_probe_det.py --frames 8 > gpurun_out/${TAG}_probe_scalar.json 2>&1
PV_RESIZE_PACKED=1 timeout 300 python scripts/gpu_probe_det.py --frames 8 > gpurun_out/${TAG}_probe_packed.json 2>&1
tail -n 1 gpurun_out/${TAG}_probe_scalar.json gpurun_out/${TAG}_probe_packed.json
