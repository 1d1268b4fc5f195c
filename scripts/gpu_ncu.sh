#!/bin/bash
# ncu evidence for profiles/ (kept small: gpurun_out/ travels back only below 64 MiB): --set full of the kernels of ONE
# serialised bench step (4 frames) exported as raw CSV on the box; the strip kernel additionally with source (.ncu-rep);
# the embedder's conv launches at batch 1024 and the tcgen05 Gram kernel as raw CSV.
TAG=${1:-r02}
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --clock-control none"
timeout 900 $NCU --set full --profile-from-start off -o /tmp/ncu/step_full -f python bench.py --frames-per-step 4 --steps 1 --warmup 3 --no-cpu-baseline --ncu-step > gpurun_out/${TAG}_ncu_step.log 2>&1
ncu -i /tmp/ncu/step_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_step_raw.csv 2>/dev/null
timeout 600 $NCU --set full --import-source on --profile-from-start off -k regex:c12_kernel -o gpurun_out/${TAG}_c12_src -f python bench.py --frames-per-step 4 --steps 1 --warmup 3 --no-cpu-baseline --ncu-step > gpurun_out/${TAG}_ncu_c12.log 2>&1
timeout 600 $NCU --set full -k regex:'srgemm_kernel|rsconv_kernel' -c 29 -o /tmp/ncu/embed_full -f python scripts/gpu_bench_aux.py --only embed --batch 1024 > gpurun_out/${TAG}_ncu_embed.log 2>&1
ncu -i /tmp/ncu/embed_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_embed_raw.csv 2>/dev/null
timeout 600 $NCU --set full -k regex:gram -c 1 -o /tmp/ncu/gram_full -f python scripts/gpu_bench_aux.py --only cluster --n 100000 > gpurun_out/${TAG}_ncu_gram.log 2>&1
ncu -i /tmp/ncu/gram_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_gram_raw.csv 2>/dev/null
du -sh gpurun_out
