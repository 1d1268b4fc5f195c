#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error|Error|rel L2|assert" gpurun_out/pytest_gpu.log | tail -25
rm -f gpurun_out/probe_srgemm.jsonl
timeout 600 python scripts/gpu_probe_srgemm.py --timing > gpurun_out/probe_timing.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench_n1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile-convs 0 > gpurun_out/bench_ncu.log 2>&1
ls -la gpurun_out
