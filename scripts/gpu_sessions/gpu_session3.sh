#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error|rel L2|max\|d\|" gpurun_out/pytest_gpu.log | tail -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
rm -f gpurun_out/probe_srgemm.jsonl
timeout 600 python scripts/gpu_probe_srgemm.py --timing > gpurun_out/probe_timing.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "bench rc=$?"; tail -4 gpurun_out/bench_n1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile-convs 0 > gpurun_out/bench_ncu.log 2>&1
for c in 25 31; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:srgemm -s 30 -c 1 -f -o gpurun_out/prof_srgemm_case$c python scripts/gpu_probe_srgemm.py --case $c > gpurun_out/ncu_case$c.log 2>&1
done
ls -la gpurun_out
