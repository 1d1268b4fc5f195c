#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error|Error|assert" gpurun_out/pytest_gpu.log | tail -6
rm -f gpurun_out/aux_bench.jsonl
timeout 900 python scripts/gpu_bench_aux.py --frames 100 > gpurun_out/aux_bench.log 2>&1; tail -3 gpurun_out/aux_bench.log | cut -c1-600
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
