#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error|Error|rel L2|assert" gpurun_out/pytest_gpu.log | tail -25
rm -f gpurun_out/probe_srgemm.jsonl
PV_SPLITS=1,2,4 timeout 900 python scripts/gpu_probe_srgemm.py --timing > gpurun_out/probe_timing.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench_n1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:srgemm -s 30 -c 1 -f -o gpurun_out/prof_srgemm_c3 python scripts/gpu_probe_srgemm.py --case 25 > gpurun_out/ncu_c3.log 2>&1
ls -la gpurun_out
