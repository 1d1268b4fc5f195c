#!/bin/bash
# state check after re-entry: full gpu tests, bench line, per-layer probe, launch list, ncu --set full of the top kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt; nproc >> gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error|Error|assert" gpurun_out/pytest_gpu.log | tail -15
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_n1.log | cut -c1-2500
timeout 300 python scripts/gpu_probe_det.py --frames 8 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --profile-convs 0 > gpurun_out/bench_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"srgemm|conv1_fused" -c 7 -o gpurun_out/r01_det_convs python scripts/gpu_probe_det.py --frames 2 --once > gpurun_out/ncu_det.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_det.log
ls -la gpurun_out | tail -12
