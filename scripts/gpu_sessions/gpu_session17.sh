#!/bin/bash
# guillotine plane packing + 4-row conv1 tiles: parity, probe, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nets_gpu.py tests/test_detconv_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x > gpurun_out/pytest_a.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|Error|assert" gpurun_out/pytest_a.log | tail -8
timeout 300 python scripts/gpu_probe_det.py --frames 8 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_n1.log | cut -c1-2500
