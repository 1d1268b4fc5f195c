#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"detconv" -c 6 -o gpurun_out/r01_detconv python scripts/gpu_probe_det.py --frames 2 --once > gpurun_out/ncu_detconv.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_detconv.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_n1.log | cut -c1-2500
