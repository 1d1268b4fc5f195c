#!/bin/bash
# final evidence of the round: ncu --set full of the conv kernels (traffic), launch list of bench steps, bench line
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rsconv|conv1_fused" -c 7 -o gpurun_out/r01_convs_v3 python scripts/gpu_probe_det.py --frames 2 --once > gpurun_out/ncu_convs.log 2>&1; echo "ncu full rc=$?"; tail -1 gpurun_out/ncu_convs.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 130 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --profile-convs 0 > gpurun_out/bench_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_n1.log | cut -c1-3500
