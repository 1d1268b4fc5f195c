#!/bin/bash
# GPU session: parity tests, smoke, srgemm timing probe, ncu captures.
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
rm -f gpurun_out/probe_srgemm.jsonl
timeout 600 python scripts/gpu_probe_srgemm.py --timing > gpurun_out/probe_timing.log 2>&1
for c in 25 31; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:srgemm -s 30 -c 2 -f -o gpurun_out/prof_srgemm_case$c python scripts/gpu_probe_srgemm.py --case $c > gpurun_out/ncu_case$c.log 2>&1
done
ls -la gpurun_out
