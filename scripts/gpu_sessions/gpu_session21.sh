#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"detconv|conv1_fused" -c 7 -o gpurun_out/r01_convs_v2 python scripts/gpu_probe_det.py --frames 2 --once > gpurun_out/ncu_convs.log 2>&1; echo "ncu full rc=$?"; tail -1 gpurun_out/ncu_convs.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --frames-per-step 16 > gpurun_out/bench_n1_b16.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_n1_b16.log | cut -c 1-3000
