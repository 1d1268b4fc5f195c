#!/bin/bash
mkdir -p gpurun_out
for mode in "" "--no-pipeline"; do
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --frames-per-step 16 $mode > gpurun_out/bench_p$mode.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_p$mode.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','host_enqueue_ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['all_convs'], d['roofline']['stage_ms_per_step'])" || tail -5 gpurun_out/bench_p$mode.log
done
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --frames-per-step 8 > gpurun_out/bench_p8.log 2>&1; tail -1 gpurun_out/bench_p8.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])"
