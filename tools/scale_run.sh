#!/bin/bash
# Round-end multi-GPU evidence on one box: the sharded generation loop and bench.py at N = $1 GPUs.
N=${1:-8}
if [ "$N" = "8" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 \
    tools/run_generations.py --games 1024 --sims 64 --iterations 2 --min-replay 5000 --eval-games 32 > gpurun_out/gen8.log 2> gpurun_out/gen8.err
  tail -2 gpurun_out/gen8.log | cut -c1-400
fi
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29630 + N)) \
  bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -1 gpurun_out/bench_${N}gpu.json | cut -c1-200
