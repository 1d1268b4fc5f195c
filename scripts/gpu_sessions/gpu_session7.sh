#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error|Error|assert" gpurun_out/pytest_gpu.log | tail -25
rm -f gpurun_out/probe_srgemm.jsonl
PV_SPLITS=1 timeout 600 python scripts/gpu_probe_srgemm.py --timing > gpurun_out/probe_timing_auto.log 2>&1
PV_MMAW=1 PV_CTAS=1 PV_SPLITS=1 timeout 600 python scripts/gpu_probe_srgemm.py --timing > gpurun_out/probe_timing_w1c1.log 2>&1
PV_MMAW=4 PV_CTAS=1 PV_SPLITS=1 timeout 600 python scripts/gpu_probe_srgemm.py --timing > gpurun_out/probe_timing_w4c1.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench_n1.log | cut -c1-2000
PV_DET_CONV1=pixrows timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_pixrows.log 2>&1; echo "bench pixrows rc=$?"; tail -2 gpurun_out/bench_n1_pixrows.log | cut -c1-2000
ls -la gpurun_out
