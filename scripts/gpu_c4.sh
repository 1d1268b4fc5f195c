#!/bin/bash
# C4 strong scaling on the GPUs of this box: N = 1, 2, 4, 8 (as many as are visible), fixed total frame count
FRAMES=${1:-4096}
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
python scripts/bench_c4.py --frames $FRAMES > gpurun_out/c4_n1.log 2>&1; tail -n 1 gpurun_out/c4_n1.log
for N in 2 4 8; do
  if [ $N -le $NG ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) scripts/bench_c4.py --frames $FRAMES > gpurun_out/c4_n$N.log 2>&1; tail -n 1 gpurun_out/c4_n$N.log
  fi
done
if [ $NG -ge 2 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 600 gpurun_out/bench_n2.json
fi
