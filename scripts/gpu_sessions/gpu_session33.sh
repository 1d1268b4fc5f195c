#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tracker" -s 8 -c 2 -o gpurun_out/r01_tracker python scripts/gpu_bench_aux.py --frames 4 --only tracker > gpurun_out/ncu_tr.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_tr.log
