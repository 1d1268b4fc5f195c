#!/bin/bash
mkdir -p gpurun_out
python scripts/gpu_h2d_probe.py
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"resize|pyramid" --csv --log-file gpurun_out/pyr_launches.csv python scripts/gpu_probe_det.py --frames 8 --once > /dev/null 2>&1; echo "list rc=$?"
grep -E "resize|pyramid" gpurun_out/pyr_launches.csv | awk -F'","' '{print $5, $(NF)}' | head -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"resize_cols" -c 3 -o gpurun_out/r01_resize python scripts/gpu_probe_det.py --frames 8 --once > gpurun_out/ncu_resize.log 2>&1; echo "ncu rc=$?"
