#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rsconv" -c 6 -o gpurun_out/r01_rsconv python scripts/gpu_probe_det.py --frames 2 --once > gpurun_out/ncu_rs.log 2>&1; echo "ncu full rc=$?"; tail -1 gpurun_out/ncu_rs.log
