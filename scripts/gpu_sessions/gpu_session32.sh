#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tracker_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x > gpurun_out/pytest_t.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|Error|assert" gpurun_out/pytest_t.log | tail -4
rm -f gpurun_out/aux_bench.jsonl
timeout 900 python scripts/gpu_bench_aux.py --frames 100 > gpurun_out/aux_bench.log 2>&1; tail -2 gpurun_out/aux_bench.log | cut -c1-400
