#!/bin/bash
TAG=${1:-trk}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tracker_gpu.py "tests/test_fullsize_gpu.py::test_c3_256_trackers_follow_known_translation" tests/test_public_api_gpu.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 300 python scripts/gpu_bench_aux.py --only tracker --frames 200 > gpurun_out/${TAG}_aux.jsonl 2>&1; tail -n 1 gpurun_out/${TAG}_aux.jsonl
