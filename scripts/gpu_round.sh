#!/bin/bash
# One GPU-box session: parity tests, default bench (both arms), aux configs, ncu launch list.  Outputs -> gpurun_out/
TAG=${1:-r02a}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/${TAG}_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
timeout 900 python scripts/gpu_bench_aux.py --frames 200 > gpurun_out/${TAG}_aux.jsonl 2> gpurun_out/${TAG}_aux.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --ncu-step > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_bench_n1.json
