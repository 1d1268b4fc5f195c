#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nets_gpu.py -m gpu -q -x > gpurun_out/pytest_nets.log 2>&1; echo "pytest nets rc=$?"
grep -E "passed|failed|error|Error|assert" gpurun_out/pytest_nets.log | tail -15
PV_DET_CONV1=fused timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_fused.log 2>&1; echo "bench fused rc=$?"; tail -2 gpurun_out/bench_n1_fused.log | cut -c1-2000
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench_n1.log | cut -c1-1200
