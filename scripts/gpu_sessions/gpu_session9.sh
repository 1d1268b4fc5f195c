#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error|Error|assert" gpurun_out/pytest_gpu.log | tail -25
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench_n1.log | cut -c1-2200
PV_DET_CONV1=pixrows timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_pixrows.log 2>&1; echo "bench pixrows rc=$?"; tail -2 gpurun_out/bench_n1_pixrows.log | cut -c1-1500
rm -f gpurun_out/aux_bench.jsonl
timeout 900 python scripts/gpu_bench_aux.py --frames 100 > gpurun_out/aux_bench.log 2>&1; tail -3 gpurun_out/aux_bench.log | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 150 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --profile-convs 0 > gpurun_out/bench_ncu.log 2>&1
ls -la gpurun_out
