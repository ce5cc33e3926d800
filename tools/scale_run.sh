#!/bin/bash
# Round-end multi-GPU evidence on one box: bench.py at N = $1 GPUs; "gen" as second argument also runs the sharded generation loop.
N=${1:-8}
if [ "$2" = "gen" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 \
    tools/run_generations.py --games 1024 --sims 64 --iterations 2 --min-replay 5000 --eval-games 32 > gpurun_out/gen$N.log 2> gpurun_out/gen$N.err
  tail -2 gpurun_out/gen$N.log | cut -c1-400
fi
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29630 + N)) \
  bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -1 gpurun_out/bench_${N}gpu.json | cut -c1-200
