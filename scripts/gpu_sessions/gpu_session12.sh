#!/bin/bash
mkdir -p gpurun_out
PV_DET_CONV1=fused timeout 300 python scripts/gpu_probe_det.py --frames 8 2>&1 | tail -2
PV_DET_CONV1=gathered timeout 300 python scripts/gpu_probe_det.py --frames 8 2>&1 | tail -2
PV_DET_CONV1=fused timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1_fused -c 1 -o gpurun_out/r01_conv1_fused python scripts/gpu_probe_det.py --frames 2 --once > gpurun_out/ncu_c1.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_c1.log
ls -la gpurun_out | tail -5
