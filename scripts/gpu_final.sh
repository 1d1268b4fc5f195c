#!/bin/bash
TAG=${1:-fin}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tracker_gpu.py "tests/test_fullsize_gpu.py::test_c3_256_trackers_follow_known_translation" tests/test_public_api_gpu.py tests/test_edge_cases_gpu.py -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
tail -c 1500 gpurun_out/${TAG}_bench_n1.json
