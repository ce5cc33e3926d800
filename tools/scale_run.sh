#!/bin/bash
# Round-end multi-GPU evidence on one 8-GPU box: 2-GPU tests, the sharded generation loop and bench.py at N = 4 and 8.
timeout 600 python -m pytest tests/test_multi_device_gpu.py -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 \
  tools/run_generations.py --games 1024 --sims 64 --iterations 2 --min-replay 5000 --eval-games 32 > gpurun_out/gen8.log 2> gpurun_out/gen8.err
tail -2 gpurun_out/gen8.log | cut -c1-600
for n in 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29630 + n)) \
    bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
  tail -1 gpurun_out/bench_${n}gpu.json | cut -c1-200
done
