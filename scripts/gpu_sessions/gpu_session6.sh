#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error|Error|assert" gpurun_out/pytest_gpu.log | tail -25
rm -f gpurun_out/probe_srgemm.jsonl
PV_CTAS=0 PV_SPLITS=1 timeout 600 python scripts/gpu_probe_srgemm.py --timing > gpurun_out/probe_timing_auto.log 2>&1
PV_CTAS=1 PV_SPLITS=1 timeout 600 python scripts/gpu_probe_srgemm.py --timing > gpurun_out/probe_timing_cta1.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench_n1.log
ls -la gpurun_out
