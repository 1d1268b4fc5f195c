#!/bin/bash
mkdir -p gpurun_out
for sw in 0 1; do
  PV_C1_SWAP=$sw timeout 600 python -m pytest tests/test_nets_gpu.py -m gpu -q -x -k "fused" > gpurun_out/pytest_fused_$sw.log 2>&1; rc=$?
  echo "swap=$sw pytest rc=$rc"; grep -E "passed|failed|Error|assert|mismatch" gpurun_out/pytest_fused_$sw.log | tail -6
  if [ $rc -eq 0 ]; then
    PV_C1_SWAP=$sw PV_DET_CONV1=fused timeout 300 python scripts/gpu_probe_det.py --frames 8 2>&1 | tail -1
    break
  fi
done
