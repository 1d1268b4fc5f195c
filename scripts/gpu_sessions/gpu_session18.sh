#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nets_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -k "pyramid or detector" > gpurun_out/pytest_a.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|Error|assert" gpurun_out/pytest_a.log | tail -8
timeout 300 python scripts/gpu_probe_det.py --frames 8 2>&1 | tail -1
for fps in 8 16; do
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --frames-per-step $fps > gpurun_out/bench_n1_b$fps.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_n1_b$fps.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','host_enqueue_ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['stage_ms_per_step'])"
done
